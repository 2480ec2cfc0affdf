"""Import alias: `import dcue_b200` == the package in ./amplifai-deepcontentrecommenders_b200
(whose directory name is not a Python identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("amplifai-deepcontentrecommenders_b200")
