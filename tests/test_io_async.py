"""fenics_b200.io.AsyncWriter: ring / writer-thread logic on CPU tensors; snapshot consistency on the device."""
import os

import numpy as np
import pytest
import torch

from fenics_b200.io import AsyncWriter, load_array


def test_ring_backpressure_and_snapshot_semantics_cpu(tmp_path):
    t = torch.zeros(1000, dtype=torch.float64)
    with AsyncWriter(depth=2) as w:
        for k in range(7):  # more saves than slots: save() must wait for the writer, never drop or reorder
            t.fill_(float(k))
            w.save(t, tmp_path / ("s%d.npy" % k))
            t.fill_(-1.0)    # overwriting right after save() must not change what gets written
        w.flush()
        assert len(w.written) == 7
    for k in range(7):
        assert np.all(load_array(str(tmp_path / ("s%d.npy" % k))) == float(k))
    with AsyncWriter(depth=1) as w:
        w.save(t, tmp_path / "missing_dir" / "x.npy")
        with pytest.raises(OSError):
            w.flush()
    # .h5 = dolfin's HDF5File layout (/f/vector_0 ...): needs h5py; without it the name is refused, loudly
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is None:
        w = AsyncWriter()
        w.save(t, tmp_path / "u.h5")
        with pytest.raises(RuntimeError, match="h5py"):
            w.flush()
        w.close()
        with pytest.raises(RuntimeError, match="h5py"):
            load_array(str(tmp_path / "u.h5"))
    else:
        with AsyncWriter() as w:
            w.save(t, tmp_path / "u.h5")
        assert np.all(load_array(str(tmp_path / "u.h5")) == -1.0)
        with h5py.File(str(tmp_path / "u.h5"), "r") as f:
            assert "vector_0" in f["f"]


@pytest.mark.gpu
def test_device_snapshots_while_the_loop_keeps_running(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fenics_b200 import mesh as pmesh
    from fenics_b200.drivers import ldc
    drv = ldc.CompLDC(pmesh.Skew_Mesh(24, 1, 1))
    drv.t = 0.45
    drv.pressure_method = "cg"
    drv.check = False
    ref = []
    with AsyncWriter(depth=2) as w:
        for k in range(5):
            drv.step()
            w.save(drv.w1, tmp_path / ("w1-%d.npy" % k))          # no host sync here
            w.save(drv.tau1_vec, tmp_path / ("tau-%d.npy" % k))
            ref.append((drv.w1.vector().t.clone(), drv.tau1_vec.vector().t.clone()))  # stream-ordered device copies
    for k, (rw, rt) in enumerate(ref):
        assert np.array_equal(load_array(str(tmp_path / ("w1-%d.npy" % k))), rw.cpu().numpy())
        assert np.array_equal(load_array(str(tmp_path / ("tau-%d.npy" % k))), rt.cpu().numpy())
    assert not np.array_equal(ref[0][0].cpu().numpy(), ref[4][0].cpu().numpy())
