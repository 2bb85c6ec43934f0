"""CPU: the `.bin` reader / sweep merge (paddle3d_b200/io.py) against the golden produced by the
reference's own LoadPointCloud.__call__ (tests/golden/make_golden.py: AST-extracted from reader.py:91-167)."""
import os

import numpy as np
import pytest

from conftest import golden
from paddle3d_b200 import io as p3d_io


def _write(tmp_path, g):
    paths = []
    for i in range(4):
        p = os.path.join(tmp_path, "c%d.bin" % i)
        g["cloud%d" % i].tofile(p)
        paths.append(p)
    mats = [g["mat0"], None, g["mat2"]]
    sweeps = [p3d_io.Sweep(paths[i + 1], mats[i], float(g["lags"][i])) for i in range(3)]
    return paths, sweeps


def test_sweep_merge_matches_reference_bit_for_bit(tmp_path):
    g = golden("sweeps.npz")
    paths, sweeps = _write(str(tmp_path), g)
    out = p3d_io.load_point_cloud(paths[0], 5, use_dim=[0, 1, 2, 4], use_time_lag=True, sweeps=sweeps,
                                  sweep_remove_radius=1, order=g["order"])
    assert out.dtype == np.float32 and out.shape == g["merged"].shape
    assert np.array_equal(out, g["merged"])
    # a different sweep order permutes blocks of rows, nothing else
    other = p3d_io.load_point_cloud(paths[0], 5, use_dim=[0, 1, 2, 4], use_time_lag=True, sweeps=sweeps,
                                    sweep_remove_radius=1)
    assert other.shape == out.shape
    key = lambda a: a[np.lexsort(a.T[::-1])]  # noqa: E731
    assert np.array_equal(key(other), key(out))


def test_reader_edge_cases(tmp_path):
    g = golden("sweeps.npz")
    paths, sweeps = _write(str(tmp_path), g)
    k = p3d_io.read_bin(paths[0], 5, use_dim=4)
    assert k.shape == (500, 4) and np.array_equal(k, g["cloud0"][:, :4])
    plain = p3d_io.load_point_cloud(paths[0], 5)  # no sweeps, no lag: the file as it is
    assert np.array_equal(plain, g["cloud0"])
    lag = p3d_io.load_point_cloud(paths[0], 5, use_dim=[0, 1, 2, 3], use_time_lag=True)
    assert lag.shape == (500, 5) and (lag[:, 4] == 0).all()
    # every sweep point inside the removal square is dropped, the key frame keeps its own
    far = p3d_io.load_point_cloud(paths[0], 5, use_dim=[0, 1, 2], sweeps=sweeps[1:2], sweep_remove_radius=100.0)
    assert far.shape == (500, 3)
    empty = os.path.join(str(tmp_path), "empty.bin")
    np.zeros((0, 5), np.float32).tofile(empty)
    assert p3d_io.load_point_cloud(empty, 5, use_dim=3).shape == (0, 3)
    merged = p3d_io.load_point_cloud(empty, 5, use_dim=[0, 1, 2], sweeps=[p3d_io.Sweep(empty)], use_time_lag=True)
    assert merged.shape == (0, 4)
    with pytest.raises(ValueError):
        p3d_io.load_point_cloud(paths[0], 5, sweeps=sweeps, order=[0, 0, 1])
    with pytest.raises(ValueError):
        p3d_io.read_bin(paths[0], 7)  # 2500 floats are not a multiple of 7


def test_deploy_preprocess_and_result_format(tmp_path):
    """deploy.preprocess / parse_result against deploy/centerpoint/python/infer.py:86-104,139-160: column selection, the
    zero time-lag column, the exact text of a 9-value and a 7-value detection line, fake rows (score -1) skipped."""
    from paddle3d_b200 import deploy
    rng = np.random.default_rng(0)
    raw = rng.normal(size=(50, 5)).astype(np.float32)
    f = tmp_path / "s.bin"
    raw.tofile(f)
    p = deploy.preprocess(str(f), 5, True)
    assert p.shape == (50, 5) and np.array_equal(p[:, :4], raw[:, :4]) and not p[:, 4].any()
    assert deploy.preprocess(str(f), 5, False).shape == (50, 4)
    boxes = np.arange(18, dtype=np.float32).reshape(2, 9)
    lines = deploy.format_result(boxes, np.array([3, 0]), np.array([0.5, -1.0], np.float32))
    assert lines == ["Score: 0.5 Label: 3 Box(x_c, y_c, z_c, w, l, h, vec_x, vec_y, -rot): 0.0 1.0 2.0 3.0 4.0 5.0 6.0 7.0 8.0"]
    l7 = deploy.format_result(boxes[:, :7], np.array([1, 2]), np.array([0.25, 0.75], np.float32))
    assert l7[1] == "Score: 0.75 Label: 2 Box(x_c, y_c, z_c, w, l, h, -rot): 9.0 10.0 11.0 12.0 13.0 14.0 15.0"
    out = tmp_path / "r.txt"
    deploy.write_results(str(out), boxes, np.array([3, 0]), np.array([0.5, 0.9], np.float32))
    assert len(out.read_text().splitlines()) == 2
