"""Importable alias of the ``ann-solo_b200/`` directory (a hyphen cannot appear in a Python
package name). All code lives in ``ann-solo_b200/``; this file only redirects the package path."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ann-solo_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
