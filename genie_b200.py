"""Import alias: `import genie_b200` == importlib.import_module("1xgpt_b200") (the package directory name
starts with a digit, so it cannot appear in an `import` statement)."""
import importlib
import sys

_pkg = importlib.import_module("1xgpt_b200")
sys.modules[__name__] = _pkg
