"""Generate tests/golden/ref_py_pipeline.npz from the REFERENCE's own Python functions, in the build container
(the modules themselves cannot be imported: they import h5py / plyfile at the top, absent here -- so the function
definitions are lifted out of the source files by name with `ast` and executed with numpy only; no reference source is
copied into this repo).

    python tests/golden/make_golden_pipeline.py            # needs /root/reference

Functions: data_utils.py center_data, normalize_data; pointnet2/utils/provider.py rotate_point_cloud,
jitter_point_cloud, shift_point_cloud, random_scale_point_cloud, random_point_dropout.  The random draws the functions
make from np.random are reproduced by re-seeding and drawing in the same order, and stored next to the outputs, so that
tests/test_golden_pipeline.py can feed them to the restatement (oracle/pipeline_oracle.py) as arguments."""
import ast
import os
import sys

import numpy as np

REF = os.environ.get("PSA_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def lift(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def raw(b, n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((b, n, 3)) * np.array([1.5, 0.7, 1.1]) + np.array([0.3, -2.0, 0.9])).astype(np.float32)


def main():
    du = lift(os.path.join(REF, "data_utils.py"), ["center_data", "normalize_data"])
    pv = lift(os.path.join(REF, "pointnet2", "utils", "provider.py"),
              ["rotate_point_cloud", "jitter_point_cloud", "shift_point_cloud", "random_scale_point_cloud", "random_point_dropout"])
    out = {}
    B, N = 6, 512
    x = raw(B, N, 11)
    cn = du["normalize_data"](du["center_data"](x.copy()))
    out["center_normalize"] = np.asarray(cn, np.float32)
    # rotate: one np.random.uniform() per cloud, in order
    np.random.seed(100)
    rot = pv["rotate_point_cloud"](cn)
    np.random.seed(100)
    out["angles"] = np.array([np.random.uniform() * 2 * np.pi for _ in range(B)])
    out["rotated"] = rot
    # jitter: one randn(B,N,3)
    np.random.seed(101)
    jit = pv["jitter_point_cloud"](rot)
    np.random.seed(101)
    out["noise"] = np.random.randn(B, N, 3)
    out["jittered_f64"] = jit                                   # float64 array, fed to a float32 placeholder by train.py
    # scale, shift, dropout (provider functions the in-scope scripts do not call; order as in pipeline_oracle.augment)
    np.random.seed(102)
    sc = pv["random_scale_point_cloud"](cn.copy())
    np.random.seed(102)
    out["scales"] = np.random.uniform(0.8, 1.25, B)
    out["scaled"] = sc
    np.random.seed(103)
    sh = pv["shift_point_cloud"](cn.copy())
    np.random.seed(103)
    out["shifts"] = np.random.uniform(-0.1, 0.1, (B, 3))
    out["shifted"] = sh
    np.random.seed(104)
    dr = pv["random_point_dropout"](cn.copy())
    np.random.seed(104)
    mask = np.zeros((B, N), bool)
    for b in range(B):
        ratio = np.random.random() * 0.875
        mask[b] = np.random.random((N)) <= ratio
    out["drop"] = mask
    out["dropped"] = dr
    dst = os.path.join(ROOT, "tests", "golden", "ref_py_pipeline.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
