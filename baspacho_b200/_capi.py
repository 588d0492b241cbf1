"""ctypes binding of the C ABI in include/baspacho_b200.h.

The same binding class serves the product library (prefix ``bspb200_``, device pointers) and - from
tests/ and bench.py only - the CPU checker library in oracle/ (prefix ``oracle_``, host pointers).
"""
import ctypes as C
import numpy as np

c_i64 = C.c_int64
c_i64p = C.POINTER(C.c_int64)
c_dblp = C.POINTER(C.c_double)
vp = C.c_void_p

BACKEND_REF, BACKEND_FAST, BACKEND_CUDA, BACKEND_SYMBOLIC_ONLY = 0, 1, 2, 100
FILL_COMPLETE, FILL_FOR_AUTO_ELIMS, FILL_FOR_GIVEN_ELIMS, FILL_NONE = 0, 1, 2, 3
MODEL_AUTO, MODEL_OPENBLAS_I7, MODEL_CUDA_2080TI, MODEL_B200 = -1, 0, 1, 2
F64, F32 = 0, 1
SOLVE_LLT, SOLVE_L, SOLVE_LT = 0, 1, 2

ARRAY_IDS = {
    "spanStart": 0, "spanToLump": 1, "lumpStart": 2, "lumpToSpan": 3, "spanOffsetInLump": 4,
    "chainColPtr": 5, "chainRowSpan": 6, "chainData": 7, "chainRowsTillEnd": 8, "boardColPtr": 9,
    "boardRowLump": 10, "boardChainColOrd": 11, "boardRowPtr": 12, "boardColLump": 13, "boardColOrd": 14,
    "permutation": 15, "sparseElimRanges": 16,
}

# every symbol include/baspacho_b200.h declares (without prefix); tests check that the .so exports all of them
DECLARED_SYMBOLS = [
    "last_error", "version", "create_solver", "create_solver_from_skel", "destroy_solver", "solver_query",
    "solver_array", "densify", "damp", "block_offset", "work_estimate", "set_stream", "set_fused", "factor",
    "factor_batched", "solve", "solve_batched", "add_mv_from", "pseudo_factor_from", "do_elimination",
    "factor_solve_host", "host_copy_bytes", "factor_solve_host_batched", "device_accessor", "dev_gemm_nt", "dev_potrf", "lumpchol_job_list", "profile_enable", "profile_report", "debug_read", "launch_count", "gen_pattern", "pattern_order", "pattern_nnz", "pattern_copy",
    "pattern_free", "random_data", "fill_reducing_permutation",
]


class BaspachoError(RuntimeError):
    pass


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _p64(a):
    return a.ctypes.data_as(c_i64p)


class CApi:
    def __init__(self, libpath, prefix):
        self.lib = C.CDLL(libpath)
        self.prefix = prefix
        self.path = libpath
        f = self._fn
        f("last_error", C.c_char_p, [])
        f("version", C.c_char_p, [])
        f("create_solver", C.c_int, [C.c_int] * 5 + [c_i64, c_i64p, c_i64p, c_i64p, c_i64, c_i64p, c_i64, c_i64p, C.POINTER(vp)])
        f("create_solver_from_skel", C.c_int, [C.c_int, C.c_int, c_i64, c_i64p, c_i64, c_i64p, c_i64p, c_i64p, c_i64, c_i64p, c_i64p, C.POINTER(vp)])
        f("destroy_solver", None, [vp])
        f("solver_query", c_i64, [vp, C.c_int])
        f("solver_array", c_i64, [vp, C.c_int, c_i64p, c_i64])
        f("densify", C.c_int, [vp, C.c_int, vp, vp, C.c_int, c_i64])
        f("damp", C.c_int, [vp, C.c_int, vp, C.c_double, C.c_double])
        f("block_offset", C.c_int, [vp, c_i64, c_i64, c_i64p, c_i64p, C.POINTER(C.c_int)])
        f("work_estimate", C.c_int, [vp, c_dblp, c_dblp, c_dblp, c_dblp, c_dblp])
        f("set_stream", C.c_int, [vp, vp])
        f("set_fused", C.c_int, [vp, C.c_int])
        f("factor", C.c_int, [vp, C.c_int, vp, c_i64, c_i64])
        f("factor_batched", C.c_int, [vp, C.c_int, C.POINTER(vp), C.c_int, c_i64, c_i64])
        f("solve", C.c_int, [vp, C.c_int, C.c_int, vp, vp, c_i64, C.c_int, c_i64, c_i64])
        f("solve_batched", C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_int, c_i64, C.c_int, c_i64, c_i64])
        f("add_mv_from", C.c_int, [vp, C.c_int, vp, c_i64, vp, c_i64, vp, c_i64, C.c_int, C.c_double])
        f("pseudo_factor_from", C.c_int, [vp, C.c_int, vp, c_i64])
        f("do_elimination", C.c_int, [vp, C.c_int, vp, C.c_int])
        f("factor_solve_host", C.c_int, [vp, C.c_int, vp, vp, vp, c_i64, C.c_int])
        if prefix == "bspb200_":
            f("dev_gemm_nt", C.c_int, [C.c_int, c_i64, c_i64, c_i64, C.c_double, vp, c_i64, vp, c_i64, C.c_double, vp, c_i64, C.c_int, vp])
            f("dev_potrf", C.c_int, [C.c_int, c_i64, c_i64, vp, c_i64, vp])
            f("lumpchol_job_list", c_i64, [C.c_int, C.c_int, C.c_int, C.c_int, vp, c_i64])
            f("device_accessor", C.c_int, [vp, C.POINTER(vp)])
            f("host_copy_bytes", C.c_int, [vp, c_i64p, c_i64p])
            f("factor_solve_host_batched", C.c_int, [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), c_i64, C.c_int])
            f("profile_enable", C.c_int, [C.c_int])
            f("profile_report", c_i64, [C.c_char_p, c_i64])
            f("debug_read", c_i64, [C.c_int, vp, c_i64])
        f("launch_count", c_i64, [])
        f("gen_pattern", C.c_int, [C.c_int, c_dblp, C.c_int, c_i64, c_i64, c_i64, C.POINTER(vp)])
        f("pattern_order", c_i64, [vp])
        f("pattern_nnz", c_i64, [vp])
        f("pattern_copy", C.c_int, [vp, c_i64p, c_i64p, c_i64p])
        f("pattern_free", None, [vp])
        f("random_data", C.c_int, [C.c_int, c_i64, C.c_double, C.c_double, c_i64, vp])
        f("fill_reducing_permutation", C.c_int, [c_i64, c_i64p, c_i64p, c_i64p])

    def _fn(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, name, fn)

    def profile(self, on):
        self.check(self.profile_enable(int(on)))

    def profile_json(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        if self.profile_report(buf, len(buf)) < 0:
            raise BaspachoError(self.last_error().decode())
        return json.loads(buf.value.decode())

    def check(self, rc):
        if rc != 0:
            raise BaspachoError(self.last_error().decode())

    # ---- synthetic problems -------------------------------------------------------------------
    def gen_pattern_arrays(self, kind, params, bsize_min, bsize_max, seed=37):
        """returns (paramSizes, ptrs, inds): CSR lower-triangular block pattern"""
        pr = (C.c_double * len(params))(*[float(x) for x in params])
        h = vp()
        self.check(self.gen_pattern(kind, pr, len(params), bsize_min, bsize_max, seed, C.byref(h)))
        try:
            n, nnz = self.pattern_order(h), self.pattern_nnz(h)
            sizes, ptrs, inds = np.empty(n, np.int64), np.empty(n + 1, np.int64), np.empty(nnz, np.int64)
            self.check(self.pattern_copy(h, _p64(sizes), _p64(ptrs), _p64(inds)))
        finally:
            self.pattern_free(h)
        return sizes, ptrs, inds

    def random_data_array(self, size, low, high, seed, dtype=np.float64):
        out = np.empty(size, dtype=dtype)
        self.check(self.random_data(F64 if dtype == np.float64 else F32, size, low, high, seed, out.ctypes.data_as(vp)))
        return out

    def amd(self, ptrs, inds):
        ptrs, inds = _i64(ptrs), _i64(inds)
        perm = np.empty(len(ptrs) - 1, np.int64)
        self.check(self.fill_reducing_permutation(len(perm), _p64(ptrs), _p64(inds), _p64(perm)))
        return perm


def dtype_code(np_dtype):
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float64:
        return F64
    if np_dtype == np.float32:
        return F32
    raise TypeError(f"unsupported dtype {np_dtype}")


class SolverHandle:
    """Owns a C-side Solver; everything that is pure host logic (skeleton queries, densify, damp, accessor)."""

    def __init__(self, api, handle):
        self.api = api
        self._h = handle
        self._arrays = {}

    def __del__(self):
        if getattr(self, "_h", None):
            self.api.destroy_solver(self._h)
            self._h = None

    @classmethod
    def create(cls, api, param_sizes, ss_ptrs, ss_inds, sparse_elim_ranges=(), elim_last_ids=(), *, backend,
               num_threads=16, find_sparse_elim_ranges=True, add_fill_policy=FILL_COMPLETE, computation_model=MODEL_AUTO):
        sizes, ptrs, inds = _i64(param_sizes), _i64(ss_ptrs), _i64(ss_inds)
        ranges, last = _i64(list(sparse_elim_ranges)), _i64(list(elim_last_ids))
        h = vp()
        api.check(api.create_solver(backend, num_threads, int(find_sparse_elim_ranges), add_fill_policy, computation_model,
                                    len(sizes), _p64(sizes), _p64(ptrs), _p64(inds), len(ranges), _p64(ranges),
                                    len(last), _p64(last), C.byref(h)))
        return cls(api, h)

    @classmethod
    def from_skel(cls, api, span_start, lump_to_span, col_ptr, row_ind, sparse_elim_ranges=(), permutation=None, *,
                  backend, num_threads=16):
        ss, lts, cp, ri = _i64(span_start), _i64(lump_to_span), _i64(col_ptr), _i64(row_ind)
        ranges = _i64(list(sparse_elim_ranges))
        perm = _i64(permutation) if permutation is not None else None
        h = vp()
        api.check(api.create_solver_from_skel(backend, num_threads, len(ss) - 1, _p64(ss), len(lts) - 1, _p64(lts), _p64(cp),
                                              _p64(ri), len(ranges), _p64(ranges), _p64(perm) if perm is not None else None,
                                              C.byref(h)))
        return cls(api, h)

    # ---- queries
    def _q(self, what):
        return int(self.api.solver_query(self._h, what))

    order = property(lambda self: self._q(0))
    data_size = property(lambda self: self._q(1))
    num_spans = property(lambda self: self._q(2))
    num_lumps = property(lambda self: self._q(3))
    can_factor_up_to_span = property(lambda self: self._q(4))
    elim_temp_size = property(lambda self: self._q(5))
    num_elim_ranges = property(lambda self: self._q(6))

    def array(self, name):
        if name not in self._arrays:
            which = ARRAY_IDS[name]
            n = self.api.solver_array(self._h, which, None, 0)
            if n < 0:
                raise BaspachoError(self.api.last_error().decode())
            out = np.empty(n, np.int64)
            self.api.solver_array(self._h, which, _p64(out), n)
            self._arrays[name] = out
        return self._arrays[name]

    def __getattr__(self, name):  # skel arrays as attributes: solver.chainData etc.
        if name in ARRAY_IDS:
            return self.array(name)
        raise AttributeError(name)

    def span_vector_offset(self, span):
        return int(self.spanStart[span])

    def span_matrix_offset(self, span):
        assert self.spanOffsetInLump[span] == 0
        return int(self.chainData[self.chainColPtr[self.spanToLump[span]]])

    def densify(self, data, fill_upper_half=False, start_span=0):
        data = np.ascontiguousarray(data)
        n = self.order - int(self.spanStart[start_span])
        dense = np.empty((n, n), dtype=data.dtype)
        self.api.check(self.api.densify(self._h, dtype_code(data.dtype), data.ctypes.data_as(vp), dense.ctypes.data_as(vp),
                                        int(fill_upper_half), start_span))
        return dense

    def damp(self, data, alpha, beta):
        assert data.flags.c_contiguous
        self.api.check(self.api.damp(self._h, dtype_code(data.dtype), data.ctypes.data_as(vp), alpha, beta))

    def block_offset(self, row_block, col_block):
        off, stride, flip = c_i64(), c_i64(), C.c_int()
        self.api.check(self.api.block_offset(self._h, row_block, col_block, C.byref(off), C.byref(stride), C.byref(flip)))
        return off.value, stride.value, bool(flip.value)

    def work_estimate(self):
        v = [C.c_double() for _ in range(5)]
        self.api.check(self.api.work_estimate(self._h, *[C.byref(x) for x in v]))
        keys = ["factor_flops", "solve_flops_per_rhs", "nnz_l", "elim_bytes", "elim_flops"]
        return {k: x.value for k, x in zip(keys, v)}

    def set_fused(self, enabled):
        self.api.check(self.api.set_fused(self._h, int(enabled)))

    # ---- raw-pointer numeric calls (pointer = int address; device for product, host for oracle)
    def factor_ptr(self, dtype, ptr, start_span=0, end_span=-1):
        self.api.check(self.api.factor(self._h, dtype, vp(ptr), start_span, end_span))

    def solve_ptr(self, dtype, mode, mat_ptr, vec_ptr, ld, n_rhs, start_span=0, end_span=-1):
        self.api.check(self.api.solve(self._h, dtype, mode, vp(mat_ptr), vp(vec_ptr), ld, n_rhs, start_span, end_span))

    def factor_batched_ptrs(self, dtype, ptrs, start_span=0, end_span=-1):
        arr = (vp * len(ptrs))(*[vp(p) for p in ptrs])
        self.api.check(self.api.factor_batched(self._h, dtype, arr, len(ptrs), start_span, end_span))

    def solve_batched_ptrs(self, dtype, mode, mat_ptrs, vec_ptrs, ld, n_rhs, start_span=0, end_span=-1):
        m = (vp * len(mat_ptrs))(*[vp(p) for p in mat_ptrs])
        v = (vp * len(vec_ptrs))(*[vp(p) for p in vec_ptrs])
        self.api.check(self.api.solve_batched(self._h, dtype, mode, m, v, len(mat_ptrs), ld, n_rhs, start_span, end_span))

    def add_mv_from_ptr(self, dtype, mat_ptr, span_index, in_ptr, in_stride, out_ptr, out_stride, n_rhs, alpha=1.0):
        self.api.check(self.api.add_mv_from(self._h, dtype, vp(mat_ptr), span_index, vp(in_ptr), in_stride, vp(out_ptr),
                                            out_stride, n_rhs, alpha))

    def pseudo_factor_from_ptr(self, dtype, ptr, span_index):
        self.api.check(self.api.pseudo_factor_from(self._h, dtype, vp(ptr), span_index))

    def do_elimination_ptr(self, dtype, ptr, range_index):
        self.api.check(self.api.do_elimination(self._h, dtype, vp(ptr), range_index))

    def factor_solve_host_ptr(self, dtype, data_ptr, factor_out_ptr, vec_ptr, ld, n_rhs):
        self.api.check(self.api.factor_solve_host(self._h, dtype, vp(data_ptr), vp(factor_out_ptr) if factor_out_ptr else None,
                                                  vp(vec_ptr) if vec_ptr else None, ld, n_rhs))
