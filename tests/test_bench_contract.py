"""CPU suite: bench.py's output contract.  Round 1 lost the `roofline` key to a `#` comment inside the dict literal of
the result line; this test parses bench.py and checks that every key the driver reads is a literal key of the JSON
line of BOTH arms (no GPU needed: nothing is executed)."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config"}


def _line_dicts():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) and \
                node.targets[0].id == "line" and isinstance(node.value, ast.Dict):
            out.append({k.value for k in node.value.keys if isinstance(k, ast.Constant)})
    return out


def test_both_arms_print_the_contract_keys():
    dicts = _line_dicts()
    assert len(dicts) == 2, "expected one result line per arm (reference, b200)"
    ref = next(d for d in dicts if "impl" in d)
    own = next(d for d in dicts if "impl" not in d)
    assert BASE_KEYS <= ref and BASE_KEYS <= own
    assert {"roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"} <= own, own
    assert {"impl", "cpu_baseline", "e2e"} <= ref


def test_roofline_object_has_the_required_fields():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id == "roofline" and \
                isinstance(node.value, ast.Dict):
            keys = {k.value for k in node.value.keys if isinstance(k, ast.Constant)}
            assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= keys
            return
    raise AssertionError("bench.py builds no roofline object")
