"""TEST / BASELINE INFRASTRUCTURE ONLY — times the UNMODIFIED reference (Exferro/anqs_quantum_chemistry, Python/PyTorch)
on host cores, for `bench.py --impl reference` and the `cpu_baseline` leg.  Product code never imports this module.

The reference is a pure-Python tree, so it travels: `__graft_entry__.build()` copies `/root/reference/nqs/nqs/**/*.py`
verbatim into `oracle/_ref/nqs/nqs/` (git-ignored, NOT gpurun-ignored), and this module imports it from there through
`oracle/ref_shim.py` (three stubs: cupy.RawKernel, openfermion.QubitOperator, torch.cuda.current_stream — the reference
code that runs is the reference's own).  Nothing here reads `/root/reference` at run time on the GPU box.

What is timed is the reference's own public call for the hot path:
    PauliObservable.compute_var_local_energy_proxy(unq_batch_as_base_indices, unq_batch_as_amps, coupling_method,
                                                   chunk_size, alpha_num, beta_num)          (pauli_observable.py:396-487)
on `torch.set_num_threads(os.cpu_count())`, with the six Hamiltonian table tensors handed over through the reference's own
`.npy` cache (pauli_observable.py:119-129) so its O(T) python table builder is not part of the timing.
"""
import os
import shutil
import tempfile
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIPPED_ROOT = os.path.join(_HERE, '_ref')
CONTAINER_ROOT = '/root/reference'


def ship_reference(force: bool = False) -> bool:
    """Copies the reference's python package into oracle/_ref (build container only).  Returns True when oracle/_ref holds it."""
    src = os.path.join(CONTAINER_ROOT, 'nqs', 'nqs')
    dst = os.path.join(SHIPPED_ROOT, 'nqs', 'nqs')
    if not os.path.isdir(src):
        return os.path.isdir(dst)
    if os.path.isdir(dst) and not force:
        return True
    shutil.rmtree(os.path.join(SHIPPED_ROOT, 'nqs'), ignore_errors=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    return True


def reference_root():
    """oracle/_ref when shipped (the GPU box), else the container's /root/reference, else None."""
    for root in (SHIPPED_ROOT, CONTAINER_ROOT):
        if os.path.isdir(os.path.join(root, 'nqs', 'nqs')):
            return root
    return None


def load():
    root = reference_root()
    if root is None:
        raise RuntimeError('reference tree not found (oracle/_ref is produced by __graft_entry__.build() in the build container)')
    os.environ['ANQS_REFERENCE_ROOT'] = root
    from oracle import ref_shim
    ref_shim.REFERENCE_ROOT = root
    return ref_shim.load_reference()


class ReferenceLocalEnergy:
    """The reference's HilbertSpace + PauliObservable on CPU for a Hamiltonian given as (xy, yz, w) arrays."""

    def __init__(self, xy, yz, w, qubit_num, threads=None):
        import torch
        from oracle import hamiltonian_oracle as orc
        self.ref = load()
        self.threads = int(threads or os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        self.torch = torch
        self.parent_dir = tempfile.mkdtemp(prefix='anqs_refarm_')
        tab = orc.Tables(xy, yz, w)  # the six tensors, pinned bit for bit to the reference's by tests/test_oracle_hamiltonian.py
        for name, arr in (('unq_xy_masks', tab.unq_xy_masks.reshape(-1, 1)), ('unq_xy_masks_inv', tab.unq_xy_masks_inv),
                          ('unq_xy_to_yz_num', tab.unq_xy_to_yz_num), ('unq_xy_to_yz_start', tab.unq_xy_to_yz_start),
                          ('rearranged_yz', tab.rearranged_yz.reshape(-1, 1)), ('rearranged_weights', tab.rearranged_weights)):
            np.save(os.path.join(self.parent_dir, f'{name}.npy'), arr)
        self.term_num, self.unq_xy_masks_num = int(tab.term_num), int(tab.unq_xy_masks_num)
        self.hs = self.ref.HilbertSpace(qubit_num=qubit_num, device=torch.device('cpu'), parent_dir=self.parent_dir, rng_seed=0,
                                        popcount_mode='memory_efficient')
        self.hs.init_perm()

        class _Terms:  # with the cache present the constructor only reads len(terms) and count_qubits(op) (PO:99-101)
            def __init__(self, n_terms, n):
                self.n_terms, self.n = n_terms, n

            def __len__(self):
                return self.n_terms

            def __iter__(self):
                yield ((self.n - 1, 'Z'),)

        class _Op:
            def __init__(self, n_terms, n):
                self.terms = _Terms(n_terms, n)
        self.ham = self.ref.PauliObservable(hilbert_space=self.hs, of_qubit_operator=_Op(self.term_num, qubit_num))
        assert int(self.ham.unq_xy_masks_num) == self.unq_xy_masks_num

    def __call__(self, samples, amps, alpha_num, beta_num, coupling_method='ham', chunk_size=None):
        """samples uint64/int64 [N], amps complex128 [N] -> E_loc complex128 [N] (numpy)."""
        torch = self.torch
        idx = torch.from_numpy(np.ascontiguousarray(samples).view(np.int64).copy()).view(-1, 1)
        a = torch.from_numpy(np.ascontiguousarray(amps, dtype=np.complex128).copy())
        if chunk_size is None:  # chunk x U x 8 B x ~6 live tensors within ~8 GB of host memory
            chunk_size = max(64, min(20000, int(8e9 // (self.unq_xy_masks_num * 8 * 6))))
        with torch.no_grad():
            e, _, _ = self.ham.compute_var_local_energy_proxy(unq_batch_as_base_indices=idx, unq_batch_as_amps=a,
                                                             coupling_method=coupling_method, chunk_size=chunk_size,
                                                             alpha_num=alpha_num, beta_num=beta_num)
        return e.numpy()

    def close(self):
        shutil.rmtree(self.parent_dir, ignore_errors=True)


def time_reference(xy, yz, w, qubit_num, samples, amps, alpha_num, beta_num, coupling_method='ham', steps=1, warmup=0,
                   threads=None):
    """Returns (E_loc/s, seconds per step list, threads, E_loc of the last step)."""
    arm = ReferenceLocalEnergy(xy, yz, w, qubit_num, threads)
    try:
        times, e = [], None
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            e = arm(samples, amps, alpha_num, beta_num, coupling_method)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        return samples.shape[0] * len(times) / sum(times), times, arm.threads, e
    finally:
        arm.close()
