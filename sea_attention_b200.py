"""Importable alias of the hyphen-named package directory `sea-attention_b200/`:
`import sea_attention_b200 as sea` == `importlib.import_module('sea-attention_b200')`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('sea-attention_b200')
for _name, _mod in list(sys.modules.items()):
    if _name.startswith('sea-attention_b200'):
        sys.modules['sea_attention_b200' + _name[len('sea-attention_b200'):]] = _mod
