"""SpMV timing on BenchCircuit-shaped rows (second form: 1..10 unit-coefficient terms per A / B row
drawn from the 10 most recently allocated variables, every second A row twice as long, one term per C row;
relations/examples/bench.rs:35-72), next to the DummyCircuit shape bench.py uses (1 term per row).
The generator is vectorised numpy, so 2^24 rows take seconds.  Needs a B200.
usage: python tools/spmv_probe.py [log_rows=24]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

BLS_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def bench_shaped_csr(n_rows, seed=0, n_instance=1):
    """Three CSR matrices (row_ptr u64, col u32) in BenchCircuit's shape.  Row i sees the variables allocated so far
    (3 + 3 i witnesses); columns are witness index + n_instance.  Duplicate columns within a row are kept, as
    to_matrices() keeps them only after compactify -- here they are merged by count, i.e. coefficients are small
    integers like the reference's inlined LCs produce."""
    rng = np.random.default_rng(seed)
    allocated = 3 + 3 * np.arange(n_rows, dtype=np.int64)
    cur = np.minimum(allocated, 10)
    lower = allocated - cur

    def matrix(sizes):
        row_ptr = np.zeros(n_rows + 1, dtype=np.uint64)
        row_ptr[1:] = np.cumsum(sizes, dtype=np.uint64)
        rows = np.repeat(np.arange(n_rows, dtype=np.int64), sizes)
        pick = (rng.random(rows.shape[0]) * cur[rows]).astype(np.int64)
        col = (lower[rows] + pick + n_instance).astype(np.uint32)
        return row_ptr, col

    na = rng.integers(1, 11, n_rows)
    nb = rng.integers(1, 11, n_rows)
    a_sizes = np.where(np.arange(n_rows) % 2 == 0, 2 * na, na)        # every second row carries the inlined extra LC
    A = matrix(a_sizes)
    B = matrix(nb)
    C = matrix(np.ones(n_rows, dtype=np.int64))
    return [A, B, C], int(3 + 3 * n_rows)


def main():
    import torch

    from snark_b200 import Backend

    log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n_rows = (1 << log_rows) - 2
    t0 = time.time()
    mats, n_wit = bench_shaped_csr(n_rows)
    one = np.array([(BLS_R_MONT_ONE >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
    be = Backend(curve=0)
    csr = [(rp, col, np.tile(one, len(col))) for rp, col in mats]
    print(f"generated {n_rows} rows, nnz = {[len(m[1]) for m in mats]} in {time.time() - t0:.1f} s", flush=True)
    m = be.r1cs_upload(n_rows, 1, n_wit, csr)
    z = torch.randint(0, 1 << 30, ((1 + n_wit), 8), dtype=torch.int32, device="cuda")       # any canonical-range limbs
    outs = [torch.empty((n_rows, 8), dtype=torch.int32, device="cuda") for _ in range(3)]
    torch.cuda.synchronize()          # the library runs on its own stream
    be.profile(True)
    for _ in range(5):
        be.lib.b2s_spmv(be.h, m, z.data_ptr(), 1, *[o.data_ptr() for o in outs])
    be.sync()
    rep = be.profile_report()
    cnt, ms = rep["spmv_kernel<Fr>"] if "spmv_kernel<Fr>" in rep else next(v for k, v in rep.items() if k.startswith("spmv"))
    nnz = sum(len(mm[1]) for mm in mats)
    bytes_alg = 40 * nnz + 3 * n_rows * 40
    print(f"spmv: {ms / cnt:.3f} ms per call, {nnz} nonzeros, {bytes_alg / (ms / cnt * 1e-3) / 1e9:.0f} GB/s algorithmic (40 B/nnz + 40 B/row/matrix)")


BLS_R_MONT_ONE = (1 << 256) % BLS_R

if __name__ == "__main__":
    main()
