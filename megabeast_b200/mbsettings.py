"""
Settings reader with the interface of ``/root/reference/megabeast/mbsettings.py``:
``mbsettings(settings_file)`` reads a Python-syntax text file (``docs/megabeast/settings.rst``)
and exposes every top-level name as an attribute (``stellar_model``, ``fd_model``,
``beast_base``, ...); ``mbsettings()`` with no file gives an empty object; ``verify()`` is the
same no-op hook.

The file is executed as one Python module body, so comments, continuation lines and ``import``
statements need no special handling.
"""
import types

__all__ = ["mbsettings"]


class mbsettings:
    """All of the settings for the MegaBEAST."""

    def __init__(self, settings_file=None):
        if settings_file is not None:
            self.settings_file = settings_file
            self.read()
            self.verify()

    def read(self):
        with open(self.settings_file, "r") as f:
            source = f.read()
        scope = {}
        exec(compile(source, self.settings_file, "exec"), scope)
        for key, value in scope.items():
            if key.startswith("__") or isinstance(value, types.ModuleType):
                continue
            setattr(self, key, value)

    def verify(self):
        """Hook for input validation (nothing is checked upstream either)."""
        pass
