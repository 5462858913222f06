"""CPU: the KNN oracle against the golden vectors of the real reference extension, and the .ply reader/writer
(reference attribute layout, scene/gaussian_model.py:220-232,258-275,355-395)."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import knn_oracle                     # noqa: E402
import make_golden_knn as mk                      # noqa: E402 (input generators only)

KNN_GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "knn_*.npz")))


@pytest.mark.parametrize("path", KNN_GOLD, ids=[os.path.basename(p)[4:-4] for p in KNN_GOLD])
def test_knn_oracle_matches_reference_extension(path):
    z = np.load(path)
    pts = mk.knn_inputs(str(z["name"]), int(z["P"]), int(z["seed"])).numpy()
    ours = knn_oracle.mean_dist2(pts)
    ref = z["mean_dist2"]
    fin = np.isfinite(ref)
    assert np.array_equal(fin, np.isfinite(ours))
    assert np.allclose(ours[fin], ref[fin], rtol=2e-6, atol=1e-12)


def test_knn_goldens_exist():
    assert len(KNN_GOLD) >= 4, "tests/golden/knn_*.npz missing (tools/make_golden_knn.py on the GPU box)"


def _fake_model(P, deg, seed):
    g = torch.Generator().manual_seed(seed)
    K = (deg + 1) ** 2
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(xyz=r(P, 3), features_dc=r(P, 1, 3), features_rest=r(P, K - 1, 3), opacity=r(P, 1), scaling=r(P, 3),
                rotation=r(P, 4))


def test_ply_round_trip_and_attribute_order(tmp_path):
    from s3gaussian_b200 import io_ply
    m = _fake_model(257, 3, 0)
    path = str(tmp_path / "sub" / "point_cloud.ply")
    io_ply.write_gaussian_ply(path, **m)
    head = open(path, "rb").read(4096).split(b"end_header\n")[0].decode()
    props = [l.split()[-1] for l in head.splitlines() if l.startswith("property")]
    want = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)] + \
        ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    assert props == want                                    # construct_list_of_attributes order
    assert "format binary_little_endian 1.0" in head and "element vertex 257" in head
    back = io_ply.read_gaussian_ply(path, 3)
    for k, v in m.items():
        assert back[k].shape == v.shape and torch.equal(back[k], v), k
    # channel-major storage of the SH coefficients: f_rest_0..14 are channel 0 (transpose(1,2).flatten(1))
    v = io_ply.read_ply_vertices(path)
    assert np.allclose(v["f_rest_1"], m["features_rest"][:, 1, 0].numpy())
    assert np.allclose(v["f_rest_15"], m["features_rest"][:, 0, 1].numpy())
    assert np.all(v["nx"] == 0)


def test_ply_reader_rejects_wrong_degree_and_reads_ascii(tmp_path):
    from s3gaussian_b200 import io_ply
    m = _fake_model(5, 1, 1)
    path = str(tmp_path / "a.ply")
    io_ply.write_gaussian_ply(path, **m)
    with pytest.raises(ValueError, match="SH degree"):
        io_ply.read_gaussian_ply(path, 3)
    assert torch.equal(io_ply.read_gaussian_ply(path, 1)["xyz"], m["xyz"])
    asc = str(tmp_path / "b.ply")
    with open(asc, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\nend_header\n"
                "1 2 3\n4 5 6\n")
    v = io_ply.read_ply_vertices(asc)
    assert v["y"].tolist() == [2.0, 5.0]


def test_distcuda2_rejects_cpu(built_lib):
    from s3gaussian_b200.simple_knn import distCUDA2
    with pytest.raises(RuntimeError, match="no CPU path"):
        distCUDA2(torch.zeros(4, 3))
