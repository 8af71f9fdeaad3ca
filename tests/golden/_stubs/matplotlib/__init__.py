"""Empty stand-in for `matplotlib` (absent in this image); see ../prettytable/__init__.py."""
