"""Importable alias for the package directory ``polyphonic-chord-texture-disentanglement_b200/``.

The task layout names the package directory after the reference repository; that name
contains hyphens and cannot be written in an ``import`` statement.  This shim makes
``import polydis_b200`` / ``from polydis_b200 import model`` resolve into that directory
(no code lives here).
"""
import os as _os

_REPO_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
PACKAGE_DIR = _os.path.join(_REPO_ROOT, "polyphonic-chord-texture-disentanglement_b200")
__path__.append(PACKAGE_DIR)

exec(compile(open(_os.path.join(PACKAGE_DIR, "__init__.py")).read(),
             _os.path.join(PACKAGE_DIR, "__init__.py"), "exec"))
