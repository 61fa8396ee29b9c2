"""bench.py's B200 arm draws its synthetic weights / images from bench_data.py (it must not import oracle/);
the CPU arm and the parity tests draw theirs from the oracle.  Both must be the same numbers."""
import torch

import bench_data
from oracle import stylenet_oracle as O


def test_generators_match_the_oracle_bit_for_bit():
    a, b = bench_data.net_state_dict(0), O.make_net_params(seed=0)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    a, b = bench_data.vgg_state_dict(1), O.make_vgg_params(seed=1)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    for norm in (False, True):
        assert torch.equal(bench_data.image_batch(2, 24, 40, seed=7, normalized=norm), O.make_image(2, 24, 40, seed=7, normalized=norm))


def test_product_arm_of_the_bench_does_not_import_the_oracle():
    import ast
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # the reference legs of bench.py: the CPU arm, the bounded CPU sample, and the eager-on-this-GPU yardstick (stock torch ops ->
    # cuDNN/cuBLAS; a measured bar beside the product's number, never the thing measured as the product)
    allowed = {"bench.py": {"run_reference", "cpu_baseline", "gpu_eager_reference"}, "bench_train.py": set(), "bench_data.py": set()}
    for fname, ok_funcs in allowed.items():
        tree = ast.parse(open(os.path.join(root, fname)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef):
                for sub in ast.walk(node):
                    if isinstance(sub, (ast.Import, ast.ImportFrom)):
                        names = [a.name for a in sub.names] + [getattr(sub, "module", "") or ""]
                        if any(n.split(".")[0] == "oracle" for n in names):
                            assert node.name in ok_funcs, f"{fname}:{node.name} imports oracle/"
        for node in tree.body:                                    # module level
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                names = [a.name for a in node.names] + [getattr(node, "module", "") or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), f"{fname} imports oracle/ at module level"
    # and the package itself never does
    pkg = os.path.join(root, "fast_neural_style_transfer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)
