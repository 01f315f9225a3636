"""CPU suite, part 6: PLY files of the inference pipeline (row f2) — layout of the header the reference's plyfile-based
writer produces (ref: u_net_arch/data_utils.py:52-68) and a write/read round trip."""
import numpy as np

from deep3dpointclouddenoising_b200.utils import ply


def test_write_ply_header_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((100, 3)).astype(np.float32)
    votes = rng.integers(1, 9, 100).astype(np.float32)
    normals = rng.standard_normal((100, 3)).astype(np.float32)
    path = tmp_path / "denoised.ply"
    ply.write_ply(str(path), [pts, votes, normals], ["vertex", "votes", "normal"])
    raw = path.read_bytes()
    header = raw[:raw.index(b"end_header\n")].decode("ascii").split("\n")
    assert header[:2] == ["ply", "format binary_little_endian 1.0"]
    assert header[2:7] == ["element vertex 100", "comment Generated with write_ply.py", "property float x", "property float y",
                           "property float z"]
    assert "element votes 100" in header and "property float scalar_votes" in header
    assert header[-4:-1] == ["property float nx", "property float ny", "property float nz"]
    body = raw[raw.index(b"end_header\n") + len(b"end_header\n"):]
    assert len(body) == 100 * (3 + 1 + 3) * 4 and np.array_equal(np.frombuffer(body[:1200], "<f4").reshape(100, 3), pts)
    back = ply.read_ply_ls(str(path), ["vertex", "votes", "normal"])
    assert np.array_equal(back["vertex"], pts) and np.array_equal(back["votes"][:, 0], votes)
    assert np.array_equal(back["normal"], normals)


def test_read_ascii_ply(tmp_path):
    path = tmp_path / "a.ply"
    path.write_text("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                    "end_header\n0 1 2\n3.5 4 5\n")
    got = ply.read_ply_ls(str(path), ["vertex"])["vertex"]
    assert np.array_equal(got, np.array([[0, 1, 2], [3.5, 4, 5]], np.float32))
