"""Command-line driver (gh-icp_b200/cxx/ghicp_cli.cpp): file formats and argument handling on the
CPU; the registration itself needs the GPU (tests/test_zz2_prep_gpu.py)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gh-icp_b200", "cxx", "ghicp_cli")


@pytest.fixture(scope="module")
def cli(g):
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "gh-icp_b200", "cxx"), "ghicp_cli"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return CLI


def write_pcd(path, P, binary, extra_field=False):
    n = len(P)
    fields, size, typ, cnt = ("x y z", "4 4 4", "F F F", "1 1 1") if not extra_field else ("intensity x y z", "4 4 4 4", "F F F F", "1 1 1 1")
    hdr = (f"# .PCD v0.7\nVERSION 0.7\nFIELDS {fields}\nSIZE {size}\nTYPE {typ}\nCOUNT {cnt}\nWIDTH {n}\nHEIGHT 1\n"
           f"VIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {'binary' if binary else 'ascii'}\n")
    Q = P if not extra_field else np.hstack([np.full((n, 1), 7.0, np.float32), P])
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if binary:
            f.write(np.ascontiguousarray(Q, np.float32).tobytes())
        else:
            for row in Q:
                f.write((" ".join(repr(float(v)) for v in row) + "\n").encode())


def write_ply(path, P, binary):
    n = len(P)
    hdr = (f"ply\nformat {'binary_little_endian' if binary else 'ascii'} 1.0\ncomment test\nelement vertex {n}\nproperty float x\n"
           "property float y\nproperty float z\nproperty uchar red\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        for row in P:
            if binary:
                f.write(struct.pack("<fffB", float(row[0]), float(row[1]), float(row[2]), 200))
            else:
                f.write(f"{float(row[0])!r} {float(row[1])!r} {float(row[2])!r} 200\n".encode())


def read_pcd_binary(path):
    raw = open(path, "rb").read()
    k = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    return np.frombuffer(raw[k:], np.float32).reshape(-1, 3)


@pytest.mark.parametrize("kind", ["pcd_ascii", "pcd_binary", "pcd_binary_extra", "ply_ascii", "ply_binary", "txt"])
def test_convert_round_trips_every_input_format(cli, tmp_path, kind):
    P = (np.random.default_rng(3).random((257, 3)) * [100, 50, 10] - [50, 0, 1]).astype(np.float32)
    src = str(tmp_path / ("in." + kind.split("_")[0]))
    if kind.startswith("pcd"):
        write_pcd(src, P, "binary" in kind, extra_field=kind.endswith("extra"))
    elif kind.startswith("ply"):
        write_ply(src, P, "binary" in kind)
    else:
        np.savetxt(src, P.astype(np.float64), fmt="%.9g")
    out = str(tmp_path / "out.pcd")
    r = subprocess.run([cli, "--convert", src, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(read_pcd_binary(out), P)
    # and out through the other writers
    for ext in ("ply", "txt"):
        o2 = str(tmp_path / ("o." + ext)); back = str(tmp_path / ("back_" + ext + ".pcd"))
        assert subprocess.run([cli, "--convert", out, o2]).returncode == 0
        assert subprocess.run([cli, "--convert", o2, back]).returncode == 0
        Q = read_pcd_binary(back)
        assert np.array_equal(Q, P) if ext == "ply" else np.allclose(Q, P, atol=1e-6, rtol=0)   # txt = %.6f like the reference


def test_argument_errors_and_unsupported_options(cli, tmp_path):
    P = np.random.default_rng(1).random((50, 3)).astype(np.float32)
    a = str(tmp_path / "a.txt"); np.savetxt(a, P)
    base = [cli, a, a, str(tmp_path / "r.txt")]
    tail = ["0.1", "0.5", "1.0", "1.1", "0.1", "6", "0.5", "0"]
    assert subprocess.run([cli]).returncode == 2                                             # usage
    assert subprocess.run(base + ["Q", "N"] + tail, capture_output=True).returncode == 2     # unknown feature
    assert subprocess.run(base + ["N", "X"] + tail, capture_output=True).returncode == 2     # unknown correspondence method
    r = subprocess.run(base + ["F", "K"] + tail, capture_output=True, text=True)             # PCL's FPFH estimator is not provided
    assert r.returncode == 2 and "FPFH" in r.stderr
    r = subprocess.run([cli, "--convert", str(tmp_path / "missing.pcd"), str(tmp_path / "o.txt")], capture_output=True, text=True)
    assert r.returncode == 3 and "cannot open" in r.stderr
    r = subprocess.run([cli, "--convert", a, str(tmp_path / "o.las")], capture_output=True, text=True)
    assert r.returncode == 3


def test_sample_pattern_mode_generates_the_reference_pattern(g, cli, tmp_path):
    """BSCEncoder(r, 7, true) of the C++ mirror = the reference's generator (rand() from the default seed, distinct cells, no
    pair twice): the file it writes holds the shipped default pattern, which tests/test_bsc_encoder.py pins on the
    reference's own constructor; the reading constructor takes it back, and rejects a damaged file."""
    r = subprocess.run([cli, "--sample-pattern"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "read back ok" in r.stdout and "equals the shipped default pattern" in r.stdout
    assert np.array_equal(np.loadtxt(tmp_path / "sample_pattern.txt", dtype=np.int32), g.bsc_default_pattern(7))
    assert np.array_equal(g.read_sample_pattern(str(tmp_path / "sample_pattern.txt")), g.bsc_default_pattern(7))
    (tmp_path / "bad").mkdir()
    np.savetxt(tmp_path / "bad" / "sample_pattern.txt", np.array([[1, 2], [3, 99]]), fmt="%d")
    # registration mode with feature B constructs the reading encoder only after the GPU stages; the damaged file is caught by
    # the Python reader here and by the C++ constructor on a GPU box (tests/test_zz3_bsc_gpu.py runs the good-file case)
    with pytest.raises(g.GhicpError):
        g.read_sample_pattern(str(tmp_path / "bad" / "sample_pattern.txt"))


def test_registration_mode_without_a_gpu_fails_loudly(g, cli, tmp_path):
    if g.device_count() > 0:
        pytest.skip("a GPU is visible here")
    P = np.random.default_rng(1).random((500, 3)).astype(np.float32)
    a = str(tmp_path / "a.txt"); np.savetxt(a, P)
    r = subprocess.run([cli, a, a, str(tmp_path / "r.txt"), "N", "N", "0.1", "0.5", "1.0", "1.1", "0.1", "6", "0.5", "0"],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stderr
