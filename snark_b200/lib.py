"""ctypes binding of include/b200snark.h (one-to-one; see the header for semantics and the reference
interfaces each entry point replaces)."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int32, c_uint32, c_uint64, c_void_p

import numpy as np

BLS12_381, BN254 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "libb200snark.so")


class B2SError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200snark error {code}: {msg}")
        self.code = code


class PkDesc(ctypes.Structure):
    _fields_ = [
        ("n_instance", c_uint64), ("n_witness", c_uint64), ("domain_size", c_uint64),
        ("alpha_g1", c_void_p), ("beta_g1", c_void_p), ("delta_g1", c_void_p),
        ("beta_g2", c_void_p), ("delta_g2", c_void_p),
        ("a_query", c_void_p), ("a_off", c_uint64), ("a_len", c_uint64),
        ("b_g1_query", c_void_p), ("b1_off", c_uint64), ("b1_len", c_uint64),
        ("b_g2_query", c_void_p), ("b2_off", c_uint64), ("b2_len", c_uint64),
        ("h_query", c_void_p), ("h_off", c_uint64), ("h_len", c_uint64),
        ("l_query", c_void_p), ("l_off", c_uint64), ("l_len", c_uint64),
    ]


# name -> (restype, argtypes): every symbol include/b200snark.h declares
SIGNATURES = {
    "b2s_version": (c_char_p, []),
    "b2s_ctx_create": (c_int32, [c_int32, c_int32, POINTER(c_void_p)]),
    "b2s_ctx_destroy": (None, [c_void_p]),
    "b2s_last_error": (c_char_p, [c_void_p]),
    "b2s_sizes": (c_int32, [c_void_p, POINTER(c_uint32)]),
    "b2s_launch_count": (c_uint64, [c_void_p]),
    "b2s_sync": (c_int32, [c_void_p]),
    "b2s_stream": (c_void_p, [c_void_p]),
    "b2s_ntt": (c_int32, [c_void_p, c_void_p, c_uint32, c_int32, c_int32, c_int32]),
    "b2s_msm_g1": (c_int32, [c_void_p, c_void_p, c_void_p, c_uint64, c_int32, c_int32, c_void_p]),
    "b2s_msm_g2": (c_int32, [c_void_p, c_void_p, c_void_p, c_uint64, c_int32, c_int32, c_void_p]),
    "b2s_msm_g1_partial": (c_int32, [c_void_p, c_void_p, c_void_p, c_uint64, c_int32, c_int32, c_void_p]),
    "b2s_msm_g2_partial": (c_int32, [c_void_p, c_void_p, c_void_p, c_uint64, c_int32, c_int32, c_void_p]),
    "b2s_g1_sum": (c_int32, [c_void_p, c_void_p, c_uint32, c_void_p]),
    "b2s_g2_sum": (c_int32, [c_void_p, c_void_p, c_uint32, c_void_p]),
    "b2s_r1cs_upload": (c_int32, [c_void_p, c_uint64, c_uint64, c_uint64, POINTER(c_void_p), POINTER(c_void_p),
                                  POINTER(c_void_p), POINTER(c_void_p)]),
    "b2s_r1cs_upload_lcmap": (c_int32, [c_void_p, c_uint64, c_uint64, c_uint64, POINTER(c_void_p), c_uint64, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_uint32, POINTER(c_void_p)]),
    "b2s_r1cs_free": (None, [c_void_p, c_void_p]),
    "b2s_spmv": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "b2s_witness_map": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "b2s_r1cs_domain_size": (c_uint64, [c_void_p]),
    "b2s_witness_map_sim": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_uint32, c_void_p]),
    "b2s_pk_upload": (c_int32, [c_void_p, POINTER(PkDesc), c_int32, POINTER(c_void_p)]),
    "b2s_pk_free": (None, [c_void_p, c_void_p]),
    "b2s_groth16_setup": (c_int32, [c_void_p, c_void_p, c_void_p, POINTER(c_void_p)] + [c_void_p] * 5),
    "b2s_pk_query": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_uint64]),
    "b2s_groth16_prove": (c_int32, [c_void_p] * 10),
    "b2s_groth16_prove_shard": (c_int32, [c_void_p] * 9),
    "b2s_groth16_prove_resident": (c_int32, [c_void_p] * 9),
    "b2s_groth16_prove_shard_resident": (c_int32, [c_void_p] * 8),
    "b2s_profile_enable": (c_int32, [c_void_p, c_int32]),
    "b2s_profile_report": (c_int32, [c_void_p, c_char_p, c_uint64]),
    "b2s_groth16_finish": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "b2s_serialize_g1_compressed": (c_int32, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64]),
    "b2s_serialize_g2_compressed": (c_int32, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64]),
    "b2s_proof_serialize_compressed": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64]),
    "b2s_serialize_g1_uncompressed": (c_int32, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64]),
    "b2s_serialize_g2_uncompressed": (c_int32, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64]),
    "b2s_proof_serialize_uncompressed": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64]),
    "b2s_vk_serialized_size": (c_uint64, [c_void_p, c_uint64, c_int32]),
    "b2s_vk_serialize": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_int32, c_void_p, c_uint64]),
    "b2s_pk_serialized_size": (c_uint64, [c_void_p, c_void_p, c_uint64, c_int32]),
    "b2s_pk_serialize": (c_int32, [c_void_p, c_void_p, c_void_p, c_uint64, c_int32, c_void_p, c_uint64]),
    "b2s_fixed_base_g1": (c_int32, [c_void_p, c_void_p, c_uint64, c_int32, c_int32, c_void_p]),
    "b2s_fixed_base_g2": (c_int32, [c_void_p, c_void_p, c_uint64, c_int32, c_int32, c_void_p]),
    "b2s_group_unique_id": (c_int32, [c_void_p]),
    "b2s_group_create": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, POINTER(c_void_p)]),
    "b2s_group_destroy": (None, [c_void_p]),
    "b2s_groth16_prove_group": (c_int32, [c_void_p] * 10),
    "b2s_groth16_prove_group_resident": (c_int32, [c_void_p] * 9),
    "b2s_poly_op": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_int32]),
    "b2s_poly_geom": (c_int32, [c_void_p, c_void_p, c_void_p, c_uint64, c_int32, c_void_p]),
    "b2s_poly_eval": (c_int32, [c_void_p, c_void_p, c_uint64, c_void_p, c_int32, c_void_p]),
    "b2s_field_op": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_uint64]),
    "b2s_group_op": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64]),
}

_lib = None


def load_library():
    """Load libb200snark.so and type every exported symbol.  Raises if the library is missing: there is
    no Python or CPU substitute."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise B2SError(-1, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(x):
    """Host numpy array or device torch tensor (or None / raw device address) -> (address, mem flag)."""
    if x is None:
        return None, MEM_HOST
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data, MEM_HOST
    if isinstance(x, int):
        return x, MEM_DEVICE
    assert x.is_contiguous()  # torch tensor
    if x.is_cuda:
        # The library launches on its own stream (b2s_stream).  Whatever torch queued to produce this tensor must be
        # finished before that stream reads it; conversely the caller keeps the tensor alive and calls Backend.sync()
        # before torch touches anything the library wrote (device-memory calls return without synchronising).
        import torch

        torch.cuda.current_stream(x.device).synchronize()
    return x.data_ptr(), (MEM_DEVICE if x.is_cuda else MEM_HOST)


class Backend:
    """One `b2s_ctx`: a curve bound to one GPU.  Buffers are numpy uint32/uint64 arrays (host) or
    CUDA torch tensors (device), already in the C-ABI layout (Montgomery limbs)."""

    def __init__(self, curve=BLS12_381, device=0):
        self.h = None
        self.lib = load_library()
        h = c_void_p()
        st = self.lib.b2s_ctx_create(curve, device, ctypes.byref(h))
        if st != 0:
            raise B2SError(st, "b2s_ctx_create failed" + (" (no sm_100 GPU visible)" if st == 17 else ""))
        self.h = h
        self.curve = curve
        sz = (c_uint32 * 6)()
        self.lib.b2s_sizes(self.h, sz)
        self.fr_bytes, self.fq_bytes, self.g1_bytes, self.g2_bytes, self.g1x_bytes, self.g2x_bytes = list(sz)

    def close(self):
        if self.h:
            self.lib.b2s_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise B2SError(st, self.lib.b2s_last_error(self.h).decode())

    @property
    def launches(self):
        return int(self.lib.b2s_launch_count(self.h))

    @property
    def stream(self):
        return self.lib.b2s_stream(self.h)

    def sync(self):
        self._ck(self.lib.b2s_sync(self.h))

    # ---- kernels ------------------------------------------------------------------------------
    def ntt(self, data, log_n, inverse=False, coset=False):
        p, mem = _ptr(data)
        self._ck(self.lib.b2s_ntt(self.h, p, log_n, int(inverse), int(coset), mem))
        return data

    def _msm(self, fn, out_bytes, bases, scalars, n, mont):
        pb, mem = _ptr(bases)
        ps, mem2 = _ptr(scalars)
        assert n == 0 or mem == mem2
        out = np.zeros(out_bytes // 4, dtype=np.uint32)
        self._ck(fn(self.h, pb, ps, n, int(mont), mem, out.ctypes.data))
        return out

    def msm_g1(self, bases, scalars, n, mont=True):
        return self._msm(self.lib.b2s_msm_g1, self.g1_bytes, bases, scalars, n, mont)

    def msm_g2(self, bases, scalars, n, mont=True):
        return self._msm(self.lib.b2s_msm_g2, self.g2_bytes, bases, scalars, n, mont)

    def msm_g1_partial(self, bases, scalars, n, mont=True):
        return self._msm(self.lib.b2s_msm_g1_partial, self.g1x_bytes, bases, scalars, n, mont)

    def msm_g2_partial(self, bases, scalars, n, mont=True):
        return self._msm(self.lib.b2s_msm_g2_partial, self.g2x_bytes, bases, scalars, n, mont)

    def g1_sum(self, xyzz, count):
        out = np.zeros(self.g1_bytes // 4, dtype=np.uint32)
        self._ck(self.lib.b2s_g1_sum(self.h, xyzz.ctypes.data, count, out.ctypes.data))
        return out

    def g2_sum(self, xyzz, count):
        out = np.zeros(self.g2_bytes // 4, dtype=np.uint32)
        self._ck(self.lib.b2s_g2_sum(self.h, xyzz.ctypes.data, count, out.ctypes.data))
        return out

    def fixed_base(self, group, scalars, n, mont=True, out=None):
        ps, mem = _ptr(scalars)
        nbytes = (self.g1_bytes if group == 1 else self.g2_bytes) * n
        if out is None:
            assert mem == MEM_HOST
            out = np.zeros(nbytes // 4, dtype=np.uint32)
        po, mem_o = _ptr(out)
        assert mem_o == mem
        fn = self.lib.b2s_fixed_base_g1 if group == 1 else self.lib.b2s_fixed_base_g2
        self._ck(fn(self.h, ps, n, int(mont), mem, po))
        return out

    def field_op(self, field, op, a, b):
        out = np.zeros_like(a)
        limbs = (self.fq_bytes if field == 0 else self.fr_bytes) // 4
        self._ck(self.lib.b2s_field_op(self.h, field, op, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size // limbs))
        return out

    def group_op(self, group, op, a, b, k):
        out = np.zeros_like(a)
        limbs = (self.g1_bytes if group == 1 else self.g2_bytes) // 4
        self._ck(self.lib.b2s_group_op(self.h, group, op, a.ctypes.data, b.ctypes.data, k.ctypes.data, out.ctypes.data,
                                       a.size // limbs))
        return out

    # ---- R1CS ---------------------------------------------------------------------------------
    def r1cs_upload(self, n_rows, n_instance, n_witness, csr):
        """csr: three (row_ptr uint64[n_rows+1], col uint32[nnz], coeff uint32[nnz*8]) numpy triples."""
        rp = (c_void_p * 3)(*[m[0].ctypes.data for m in csr])
        col = (c_void_p * 3)(*[m[1].ctypes.data for m in csr])
        co = (c_void_p * 3)(*[m[2].ctypes.data for m in csr])
        h = c_void_p()
        self._ck(self.lib.b2s_r1cs_upload(self.h, n_rows, n_instance, n_witness, rp, col, co, ctypes.byref(h)))
        return h

    def r1cs_upload_lcmap(self, n_rows, n_instance, n_witness, args, lc_offsets, lc_vars, lc_coeffs, pool):
        """The same handle, CSR built on the device from the constraint system's LcMap.  args: three uint64[n_rows] arrays
        of raw Variables; lc_offsets uint64[n_lcs+1]; lc_vars uint64[], lc_coeffs uint32[]; pool uint32[pool_len*8]."""
        a = (c_void_p * 3)(*[x.ctypes.data for x in args])
        h = c_void_p()
        self._ck(self.lib.b2s_r1cs_upload_lcmap(self.h, n_rows, n_instance, n_witness, a, len(lc_offsets) - 1, lc_offsets.ctypes.data,
                                                lc_vars.ctypes.data, lc_coeffs.ctypes.data, pool.ctypes.data, len(pool) // 8,
                                                ctypes.byref(h)))
        return h

    def r1cs_free(self, m):
        self.lib.b2s_r1cs_free(self.h, m)

    def domain_size(self, m):
        return int(self.lib.b2s_r1cs_domain_size(m))

    def spmv(self, m, z, n_rows):
        pz, mem = _ptr(z)
        assert mem == MEM_HOST
        outs = [np.zeros(n_rows * 8, dtype=np.uint32) for _ in range(3)]
        self._ck(self.lib.b2s_spmv(self.h, m, pz, mem, *[o.ctypes.data for o in outs]))
        return outs

    def witness_map(self, m, z):
        pz, mem = _ptr(z)
        assert mem == MEM_HOST
        h = np.zeros(self.domain_size(m) * 8, dtype=np.uint32)
        self._ck(self.lib.b2s_witness_map(self.h, m, pz, mem, h.ctypes.data))
        return h

    def witness_map_sim(self, m, z, log_ranks):
        """witness_map by the distributed schedule with 2^log_ranks virtual ranks on this GPU (test entry)."""
        pz, mem = _ptr(z)
        assert mem == MEM_HOST
        h = np.zeros(self.domain_size(m) * 8, dtype=np.uint32)
        self._ck(self.lib.b2s_witness_map_sim(self.h, m, pz, mem, log_ranks, h.ctypes.data))
        return h

    # ---- Groth16 ------------------------------------------------------------------------------
    def pk_upload(self, desc: PkDesc, mem=MEM_HOST):
        h = c_void_p()
        self._ck(self.lib.b2s_pk_upload(self.h, ctypes.byref(desc), mem, ctypes.byref(h)))
        return h

    def groth16_setup(self, m, trapdoor, n_instance):
        """trapdoor: uint32[5*8] Montgomery (tau, alpha, beta, gamma, delta) -> (pk handle, vk dict of numpy arrays)."""
        h = c_void_p()
        vk = {"alpha_g1": np.zeros(self.g1_bytes // 4, dtype=np.uint32), "beta_g2": np.zeros(self.g2_bytes // 4, dtype=np.uint32),
              "gamma_g2": np.zeros(self.g2_bytes // 4, dtype=np.uint32), "delta_g2": np.zeros(self.g2_bytes // 4, dtype=np.uint32),
              "gamma_abc_g1": np.zeros(max(n_instance, 1) * self.g1_bytes // 4, dtype=np.uint32)}
        self._ck(self.lib.b2s_groth16_setup(self.h, m, trapdoor.ctypes.data, ctypes.byref(h), vk["alpha_g1"].ctypes.data,
                                            vk["beta_g2"].ctypes.data, vk["gamma_g2"].ctypes.data, vk["delta_g2"].ctypes.data,
                                            vk["gamma_abc_g1"].ctypes.data))
        return h, vk

    def pk_query(self, pk, which, count):
        per = self.g2_bytes if which in (2, 6) else self.g1_bytes
        out = np.zeros(count * per // 4, dtype=np.uint32)
        self._ck(self.lib.b2s_pk_query(self.h, pk, which, out.ctypes.data, out.nbytes))
        return out

    def serialize_points(self, group, affine, count, compressed=True):
        per = (self.fq_bytes if group == 1 else 2 * self.fq_bytes) * (1 if compressed else 2)
        out = np.zeros(count * per, dtype=np.uint8)
        name = f"b2s_serialize_g{group}_{'compressed' if compressed else 'uncompressed'}"
        self._ck(getattr(self.lib, name)(self.h, affine.ctypes.data, count, out.ctypes.data, out.nbytes))
        return out.tobytes()

    def proof_bytes(self, a, b, c, compressed=True):
        out = np.zeros(4 * self.fq_bytes * (1 if compressed else 2), dtype=np.uint8)
        fn = self.lib.b2s_proof_serialize_compressed if compressed else self.lib.b2s_proof_serialize_uncompressed
        self._ck(fn(self.h, a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, out.nbytes))
        return out.tobytes()

    def vk_bytes(self, alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1, n_gamma_abc, compressed=True):
        """ark-groth16 VerifyingKey bytes from the HOST affine points b2s_groth16_setup returned."""
        out = np.zeros(int(self.lib.b2s_vk_serialized_size(self.h, n_gamma_abc, int(compressed))), dtype=np.uint8)
        self._ck(self.lib.b2s_vk_serialize(self.h, alpha_g1.ctypes.data, beta_g2.ctypes.data, gamma_g2.ctypes.data, delta_g2.ctypes.data,
                                           gamma_abc_g1.ctypes.data, n_gamma_abc, int(compressed), out.ctypes.data, out.nbytes))
        return out.tobytes()

    def pk_bytes(self, pk, vk_bytes, compressed=True):
        """ark-groth16 ProvingKey bytes of a device-resident full key (vk_bytes from vk_bytes())."""
        vk = np.frombuffer(vk_bytes, dtype=np.uint8)
        out = np.zeros(int(self.lib.b2s_pk_serialized_size(self.h, pk, len(vk), int(compressed))), dtype=np.uint8)
        self._ck(self.lib.b2s_pk_serialize(self.h, pk, vk.ctypes.data, len(vk), int(compressed), out.ctypes.data, out.nbytes))
        return out.tobytes()

    def pk_free(self, pk):
        self.lib.b2s_pk_free(self.h, pk)

    def _proof_bufs(self):
        return (np.zeros(self.g1_bytes // 4, dtype=np.uint32), np.zeros(self.g2_bytes // 4, dtype=np.uint32),
                np.zeros(self.g1_bytes // 4, dtype=np.uint32))

    def groth16_prove(self, pk, m, z_inst, z_wit, r, s):
        a, b, c = self._proof_bufs()
        self._ck(self.lib.b2s_groth16_prove(self.h, pk, m, _ptr(z_inst)[0], _ptr(z_wit)[0], r.ctypes.data, s.ctypes.data,
                                            a.ctypes.data, b.ctypes.data, c.ctypes.data))
        return a, b, c

    def groth16_prove_resident(self, pk, m, z_dev, r, s):
        a, b, c = self._proof_bufs()
        self._ck(self.lib.b2s_groth16_prove_resident(self.h, pk, m, _ptr(z_dev)[0], r.ctypes.data, s.ctypes.data,
                                                     a.ctypes.data, b.ctypes.data, c.ctypes.data))
        return a, b, c

    def groth16_prove_shard_resident(self, pk, m, z_dev, r, s):
        g1 = np.zeros(4 * self.g1x_bytes // 4, dtype=np.uint32)
        g2 = np.zeros(self.g2x_bytes // 4, dtype=np.uint32)
        self._ck(self.lib.b2s_groth16_prove_shard_resident(self.h, pk, m, _ptr(z_dev)[0], r.ctypes.data, s.ctypes.data,
                                                           g1.ctypes.data, g2.ctypes.data))
        return g1, g2

    def profile(self, on=True):
        self._ck(self.lib.b2s_profile_enable(self.h, int(on)))

    def profile_report(self):
        """{kernel name: (launches, total_ms)} since the last report."""
        buf = ctypes.create_string_buffer(1 << 16)
        self._ck(self.lib.b2s_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split("\t")
            out[name] = (int(cnt), float(ms))
        return out

    def groth16_prove_shard(self, pk, m, z_inst, z_wit, r, s):
        g1 = np.zeros(4 * self.g1x_bytes // 4, dtype=np.uint32)
        g2 = np.zeros(self.g2x_bytes // 4, dtype=np.uint32)
        self._ck(self.lib.b2s_groth16_prove_shard(self.h, pk, m, _ptr(z_inst)[0], _ptr(z_wit)[0], r.ctypes.data, s.ctypes.data,
                                                  g1.ctypes.data, g2.ctypes.data))
        return g1, g2

    # ---- multi-GPU group (NCCL inside the library) ----------------------------------------------
    @staticmethod
    def group_unique_id():
        """128-byte NCCL id (rank 0 draws it; the host program distributes it)."""
        buf = np.zeros(128, dtype=np.uint8)
        st = load_library().b2s_group_unique_id(buf.ctypes.data)
        if st != 0:
            raise B2SError(st, "b2s_group_unique_id failed (NCCL not loadable?)")
        return buf

    def group_create(self, uid, rank, world):
        g = c_void_p()
        uid = np.ascontiguousarray(uid, dtype=np.uint8) if uid is not None else None
        self._ck(self.lib.b2s_group_create(self.h, uid.ctypes.data if uid is not None else None, rank, world, ctypes.byref(g)))
        return g

    def group_destroy(self, g):
        self.lib.b2s_group_destroy(g)

    def groth16_prove_group(self, g, pk, m, z_inst, z_wit, r, s):
        """Collective over the group; the proof (a, b, c) is meaningful on rank 0."""
        a, b, c = self._proof_bufs()
        st = self.lib.b2s_groth16_prove_group(g, pk, m, _ptr(z_inst)[0], _ptr(z_wit)[0], r.ctypes.data, s.ctypes.data, a.ctypes.data,
                                              b.ctypes.data, c.ctypes.data)
        self._ck(st)
        return a, b, c

    def groth16_prove_group_resident(self, g, pk, m, z_dev, r, s):
        a, b, c = self._proof_bufs()
        st = self.lib.b2s_groth16_prove_group_resident(g, pk, m, _ptr(z_dev)[0], r.ctypes.data, s.ctypes.data, a.ctypes.data,
                                                       b.ctypes.data, c.ctypes.data)
        self._ck(st)
        return a, b, c

    def groth16_finish(self, pk, g1_partials, g2_partials, n_shards, r, s):
        a, b, c = self._proof_bufs()
        self._ck(self.lib.b2s_groth16_finish(self.h, pk, g1_partials.ctypes.data, g2_partials.ctypes.data, n_shards,
                                             r.ctypes.data, s.ctypes.data, a.ctypes.data, b.ctypes.data, c.ctypes.data))
        return a, b, c
