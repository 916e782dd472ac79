"""Import alias: the package directory is named `video-gen-evals_b200` (not a valid identifier), so
`import tag_b200` re-exports it."""
import importlib as _importlib
import sys as _sys

_pkg = _importlib.import_module("video-gen-evals_b200")
_sys.modules[__name__] = _pkg
