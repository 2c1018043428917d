"""`oscar` package shim: `oscar.modeling.modeling_vlbert` comes from mvp_pytorch_b200, every other submodule
(`oscar.utils`, `oscar.oscar_datasets_ml`, `oscar.run_retrieval`, ...) from the reference checkout on sys.path."""
import os
import sys

__version__ = "0.1.0"


def _extend(path, package_relpath):
    here = os.path.abspath(path[0])
    for root in list(sys.path):
        cand = os.path.abspath(os.path.join(root or ".", *package_relpath))
        if cand != here and os.path.isfile(os.path.join(cand, "__init__.py")) and cand not in path:
            path.append(cand)


_extend(__path__, ("oscar",))
