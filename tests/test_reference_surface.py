"""The drop-in surface, pinned against the reference's SOURCE (parsed with `ast`: importing it needs DGL, which is
absent): every class of /root/reference/models/conv.py exists here with the same constructor arguments and defaults,
the same forward arguments, and every attribute its __init__ assigns.  Runs only where the reference tree is mounted
(this container); the GPU box never reads /root/reference."""
import ast
import inspect
import os

import pytest
from torch import nn

REF = "/root/reference/models/conv.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")


def _ref_classes():
    tree = ast.parse(open(REF).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            fns = {f.name: f for f in node.body if isinstance(f, ast.FunctionDef)}
            out[node.name] = fns
    return out


def _args(fn):
    a = fn.args
    names = [x.arg for x in a.args][1:]                       # drop self
    defaults = [ast.literal_eval(d) for d in a.defaults]
    return names, dict(zip(names[len(names) - len(defaults):], defaults))


def _assigned_attrs(fn):
    return {t.attr for n in ast.walk(fn) if isinstance(n, ast.Assign) for t in n.targets
            if isinstance(t, ast.Attribute) and isinstance(t.value, ast.Name) and t.value.id == "self"}


def test_every_reference_class_has_the_same_signature():
    import models.conv as mine                                # the reference's import path: `from models.conv import ...`
    ref = _ref_classes()
    assert {"SIRConv", "SIREConv"} <= set(ref)
    for name, fns in ref.items():
        cls = getattr(mine, name, None) or getattr(__import__("sirgcn_b200"), name)
        assert issubclass(cls, nn.Module), name
        for meth in ("__init__", "forward"):
            names, defaults = _args(fns[meth])
            sig = inspect.signature(getattr(cls, meth))
            got = [p for p in sig.parameters.values()][1:]
            assert [p.name for p in got][:len(names)] == names, (name, meth, names, [p.name for p in got])
            for p in got[:len(names)]:
                if p.name in defaults:
                    assert p.default == defaults[p.name], (name, meth, p.name)
                else:
                    assert p.default is inspect.Parameter.empty, (name, meth, p.name)
            for p in got[len(names):]:                        # anything extra must be optional
                assert p.default is not inspect.Parameter.empty, (name, meth, p.name)


def test_every_attribute_the_reference_assigns_exists():
    import sirgcn_b200 as pkg
    ref = _ref_classes()
    built = {
        "SIRConv": pkg.SIRConv(8, 16, 4, nn.ReLU(), agg_type="sym"),
        "SIREConv": pkg.SIREConv(8, 3, 16, 4, nn.ReLU(), agg_type="mean"),
        "SIRConvBase": pkg.SIRConvBase(nn.Linear(16, 4), "sum"),
        "SIREConvBase": pkg.SIREConvBase(nn.Linear(19, 4), "max"),
    }
    for name, obj in built.items():
        for attr in _assigned_attrs(ref[name]["__init__"]):
            assert hasattr(obj, attr), (name, attr)
    # state_dict layout of the two parametrised layers (what checkpoints of the reference contain)
    assert list(built["SIRConv"].state_dict()) == ["linear_query.weight", "linear_query.bias", "linear_key.weight",
                                                   "linear_relation.weight", "linear_relation.bias"]
    assert "linear_edge.weight" in built["SIREConv"].state_dict() and "linear_edge.bias" not in built["SIREConv"].state_dict()


def test_models_package_overlays_the_reference_tree(tmp_path):
    """`PYTHONPATH=<this repo>` + the scripts' own `sys.path.append('../..')`: models.conv is ours, models.utils / norm
    are the reference's files (simulated by a stand-in tree: the real ones import dgl)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = tmp_path / "refroot"
    (ref / "models").mkdir(parents=True)
    (ref / "models" / "__init__.py").write_text("")
    (ref / "models" / "conv.py").write_text("SIRConv = 'the reference layer'\n")
    (ref / "models" / "utils.py").write_text("MLP = 'reference glue'\n")
    script = (f"import sys; sys.path.append({str(ref)!r})\n"
              "from models.conv import SIRConv, SIREConv\n"
              "from models.utils import MLP\n"
              "import sirgcn_b200\n"
              "assert SIRConv is sirgcn_b200.SIRConv and MLP == 'reference glue'\n"
              "print('overlay ok')\n")
    env = dict(os.environ, PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert out.returncode == 0 and "overlay ok" in out.stdout, out.stderr[-1500:]
