"""BAL file reader / writer: semantics of src/ReadFiles.jl:9-53."""
import numpy as np
import pytest

from conftest import small_problem


def test_reader_applies_index_shift_and_camera_permutation(tmp_path, ba):
    from bundleadjustment.jl_b200 import balio
    # one camera, one point, one observation, in FILE order r t f k1 k2
    txt = "1 1 1\n0 0 -332.65 262.09\n" + "\n".join(str(v) for v in
          [0.1, 0.2, 0.3, 1.0, 2.0, 3.0, 500.0, -1e-7, 2e-13]) + "\n4.0\n5.0\n6.0\n"
    f = tmp_path / "problem-1-1-pre.txt"
    f.write_text(txt)
    cam, pnt, pt2d, x0, ncams, npnts, nobs = balio.readfile(str(f))
    assert (ncams, npnts, nobs) == (1, 1, 1)
    assert cam.tolist() == [1] and pnt.tolist() == [1]           # 0-based file -> 1-based (ReadFiles.jl:23-24)
    assert pt2d.tolist() == [-332.65, 262.09]
    assert x0.tolist() == [4.0, 5.0, 6.0, 0.1, 0.2, 0.3, 1.0, 2.0, 3.0, -1e-7, 2e-13, 500.0]  # points, then r t k1 k2 f


@pytest.mark.parametrize("ext", [".txt", ".txt.bz2"])
def test_write_read_round_trip_is_bit_exact(tmp_path, ba, ext):
    from bundleadjustment.jl_b200 import balio
    p = small_problem(ba)
    f = tmp_path / ("problem-%d-%d-pre%s" % (p.ncams, p.npnts, ext))
    balio.write_problem(str(f), p)
    cam, pnt, pt2d, x0, ncams, npnts, nobs = balio.readfile(str(f))
    assert (ncams, npnts, nobs) == (p.ncams, p.npnts, p.nobs)
    assert np.array_equal(cam, p.cam_idx) and np.array_equal(pnt, p.pnt_idx)
    assert np.array_equal(pt2d, p.pt2d) and np.array_equal(x0, p.x0)   # every FP64 bit survives


def test_reader_rejects_truncated_file(tmp_path):
    from bundleadjustment.jl_b200 import balio
    f = tmp_path / "bad.txt"
    f.write_text("1 1 1\n0 0 1.0 2.0\n0.1\n0.2\n")
    with pytest.raises(ValueError):
        balio.readfile(str(f))


def test_oracle_residual_same_after_round_trip(tmp_path, ba, oracle):
    from bundleadjustment.jl_b200 import balio
    p = small_problem(ba)
    f = tmp_path / "problem-7-60-pre.txt.bz2"
    balio.write_problem(str(f), p)
    cam, pnt, pt2d, x0, ncams, npnts, nobs = balio.readfile(str(f))
    a = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts)
    b = oracle.cons(cam, pnt, pt2d, x0, npnts)
    assert np.array_equal(a, b)
