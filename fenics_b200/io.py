"""Asynchronous hand-off of device fields to host files.

The reference writes solutions from inside / after its loops with blocking calls: `save_solution(u, filename)`
(dolfin HDF5File, "flow between eccentrically rotating cylinders"/fenics_base.py:94-99) and the `.npy` series
(lid-driven-cavity/fenics_fem.py:456-472).  With the fields resident on the GPU a blocking device->host copy +
file write would stall the step loop; AsyncWriter takes a consistent snapshot without stalling it:

  1. a device->device copy into a staging buffer, enqueued on the COMPUTE stream (stream-ordered after the kernels
     that produced the field; microseconds), so the loop may overwrite the field immediately;
  2. the device->host copy of the staging buffer into pinned memory on a SIDE stream (copy engine, overlaps the
     next steps' kernels);
  3. a writer thread waits for that copy's event and writes the file, then returns the slot to the ring: `.npy`, or
     -- for names ending in .h5 / .hdf5 -- the layout dolfin's `HDF5File.write(u, "/f")` produces, so the reference's
     `load_solution` (fenics_base.py:102-108) can read it: group `/f` with `vector_0` (the dof vector), and for a
     Function also `cell_dofs`, `x_cell_dofs` and `cells` (the cell-dof table and its offsets).  HDF5 needs h5py; without
     it a .h5 name is an ERROR (surfaced by the next save()/flush()), never a silent change of container.

`depth` slots bound the memory; when all are in flight, save() waits for the oldest (back-pressure instead of
unbounded queues).  CPU tensors are accepted too (no streams; used by the CPU test of the ring / thread logic).
"""
import queue
import threading

import numpy as np
import torch


class AsyncWriter:
    def __init__(self, depth=4):
        self.depth = int(depth)
        self._free = queue.Queue()
        for i in range(self.depth):
            self._free.put(i)
        self._slots = [dict() for _ in range(self.depth)]
        self._jobs = queue.Queue()
        self._side = None
        self._errors = []
        self.written = []
        self._th = threading.Thread(target=self._worker, daemon=True)
        self._th.start()

    # -- public ----------------------------------------------------------------------------------------
    def save(self, field, path):
        """snapshot `field` (a Function, la.Vector or tensor) now, write it to `path` later; returns at once unless
        every slot is still in flight."""
        t = field
        meta = None
        if not torch.is_tensor(t):
            if hasattr(t, "function_space"):
                meta = t.function_space()
            if hasattr(t, "vector"):
                t = t.vector()
            t = t.t  # la.Vector -> its device tensor
        self._raise_pending()
        i = self._free.get()  # back-pressure: waits for the writer thread when the ring is full
        slot = self._slots[i]
        n = t.numel()
        if t.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=t.device)
            if slot.get("n", -1) < n or slot.get("dev") != t.device:
                slot.update(stage=torch.empty(n, dtype=t.dtype, device=t.device),
                            host=torch.empty(n, dtype=t.dtype).pin_memory(), n=n, dev=t.device)
            stage, host = slot["stage"][:n], slot["host"][:n]
            stage.copy_(t.reshape(-1))  # compute stream: ordered after the producers, before later overwrites
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self._side):
                self._side.wait_event(ready)
                host.copy_(stage, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
        else:
            if slot.get("n", -1) < n or slot.get("dev") != t.device:
                slot.update(host=torch.empty(n, dtype=t.dtype), n=n, dev=t.device)
            host = slot["host"][:n]
            host.copy_(t.reshape(-1))
            done = None
        self._jobs.put((i, done, host, str(path), meta))

    def flush(self):
        """wait until everything handed over so far is on disk; raises the first write error."""
        self._jobs.join()
        self._raise_pending()

    def close(self):
        self.flush()
        self._jobs.put(None)
        self._th.join(timeout=30)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- writer thread -----------------------------------------------------------------------------------
    def _raise_pending(self):
        if self._errors:
            raise self._errors.pop(0)

    @staticmethod
    def _write(path, arr, space=None):
        if path.endswith((".h5", ".hdf5")):
            try:
                import h5py
            except ImportError:
                raise RuntimeError("cannot write %s: HDF5 output (dolfin's /f/vector_0 layout) needs h5py, which is not "
                                   "installed -- use a .npy name" % path)
            with h5py.File(path, "w") as f:
                g = f.create_group("f")
                g.create_dataset("vector_0", data=arr)
                if space is not None and getattr(space, "cell_dofs", None) is not None:
                    cd = np.asarray(space.cell_dofs)
                    g.create_dataset("cell_dofs", data=cd.ravel().astype(np.int64))
                    g.create_dataset("x_cell_dofs", data=np.arange(0, cd.size + 1, cd.shape[1], dtype=np.int64))
                    g.create_dataset("cells", data=np.arange(cd.shape[0], dtype=np.int64))
            return
        with open(path, "wb") as f:
            np.save(f, arr)

    def _worker(self):
        while True:
            job = self._jobs.get()
            if job is None:
                self._jobs.task_done()
                return
            i, done, host, path, meta = job
            try:
                if done is not None:
                    done.synchronize()
                self._write(path, host.numpy(), meta)
                self.written.append(path)
            except Exception as e:  # surfaced by the next save()/flush()
                self._errors.append(e)
            finally:
                self._free.put(i)
                self._jobs.task_done()


def load_array(path):
    """the dof vector of a file AsyncWriter (or dolfin's HDF5File.write(u, "/f")) wrote"""
    if path.endswith((".h5", ".hdf5")):
        try:
            import h5py
        except ImportError:
            raise RuntimeError("cannot read %s: HDF5 needs h5py, which is not installed" % path)
        with h5py.File(path, "r") as f:
            node = f["f"]
            return np.asarray(node["vector_0"] if isinstance(node, h5py.Group) else node)
    with open(path, "rb") as f:
        return np.load(f)
