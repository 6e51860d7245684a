"""Oracle-only stand-in for the `beast` package (see ../README.md)."""
