/*
 * C ABI of the B200-native supernodal sparse Cholesky (libbaspacho_b200.so).
 *
 * This is the drop-in boundary one level below the C++ operator interface
 * (baspacho_b200/csrc/host/MatOps.h == reference baspacho/baspacho/MatOps.h): plain pointers and sizes,
 * no C++/torch types, no exceptions. Each entry point names the reference interface it replaces.
 * All integer arrays are int64 (the reference's index type). Numeric buffers passed to the
 * factor/solve entry points are DEVICE pointers (reference Solver.h:184-188: device pointers for the
 * CUDA backend) unless the function name ends in _host. Every function returns 0 on success, non-zero on
 * error; bspb200_last_error() then holds the message (the reference throws std::runtime_error,
 * Utils.cpp:33-37). The library fails loudly (error return) when no CUDA device is usable: there is no
 * CPU fallback in this library.
 */
#ifndef BASPACHO_B200_H_
#define BASPACHO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bspb200_solver bspb200_solver;   /* opaque: a BaSpaCho::Solver (reference Solver.h:34-180) */
typedef struct bspb200_pattern bspb200_pattern; /* opaque: a block sparsity pattern + block sizes */

/* reference Solver.h:189-218 */
enum { BSPB200_BACKEND_REF = 0, BSPB200_BACKEND_FAST = 1, BSPB200_BACKEND_CUDA = 2, BSPB200_BACKEND_SYMBOLIC_ONLY = 100 };
enum { BSPB200_FILL_COMPLETE = 0, BSPB200_FILL_FOR_AUTO_ELIMS = 1, BSPB200_FILL_FOR_GIVEN_ELIMS = 2, BSPB200_FILL_NONE = 3 };
/* computation model presets (reference ComputationModel.cpp:12-30) ; -1 = automatic as in Solver.cpp:679-683 */
enum { BSPB200_MODEL_AUTO = -1, BSPB200_MODEL_OPENBLAS_I7 = 0, BSPB200_MODEL_CUDA_2080TI = 1, BSPB200_MODEL_B200 = 2 };
enum { BSPB200_F64 = 0, BSPB200_F32 = 1 };
enum { BSPB200_SOLVE_LLT = 0, BSPB200_SOLVE_L = 1, BSPB200_SOLVE_LT = 2 };

/* integer queries (bspb200_solver_query) */
enum {
  BSPB200_Q_ORDER = 0, BSPB200_Q_DATA_SIZE = 1, BSPB200_Q_NUM_SPANS = 2, BSPB200_Q_NUM_LUMPS = 3,
  BSPB200_Q_CAN_FACTOR_UP_TO = 4, BSPB200_Q_ELIM_TEMP_SIZE = 5, BSPB200_Q_NUM_ELIM_RANGES = 6
};
/* skeleton arrays (bspb200_solver_array): members of CoalescedBlockMatrixSkel, reference CoalescedBlockMatrix.h:88-110 */
enum {
  BSPB200_A_SPAN_START = 0, BSPB200_A_SPAN_TO_LUMP = 1, BSPB200_A_LUMP_START = 2, BSPB200_A_LUMP_TO_SPAN = 3,
  BSPB200_A_SPAN_OFFSET_IN_LUMP = 4, BSPB200_A_CHAIN_COL_PTR = 5, BSPB200_A_CHAIN_ROW_SPAN = 6, BSPB200_A_CHAIN_DATA = 7,
  BSPB200_A_CHAIN_ROWS_TILL_END = 8, BSPB200_A_BOARD_COL_PTR = 9, BSPB200_A_BOARD_ROW_LUMP = 10,
  BSPB200_A_BOARD_CHAIN_COL_ORD = 11, BSPB200_A_BOARD_ROW_PTR = 12, BSPB200_A_BOARD_COL_LUMP = 13,
  BSPB200_A_BOARD_COL_ORD = 14,
  BSPB200_A_PERMUTATION = 15,        /* Solver::paramToSpan(), Solver.h:131 */
  BSPB200_A_SPARSE_ELIM_RANGES = 16  /* Solver::sparseEliminationRanges(), Solver.h:128 */
};

const char* bspb200_last_error(void);
const char* bspb200_version(void);

/* ---- analysis: replaces createSolver() (reference Solver.h:235-237, Solver.cpp:611-752).
 * ss_ptrs/ss_inds: CSR lower-triangular BLOCK pattern incl. diagonal (order = n_params). */
int bspb200_create_solver(int backend, int num_threads, int find_sparse_elim_ranges, int add_fill_policy,
                          int computation_model, int64_t n_params, const int64_t* param_sizes,
                          const int64_t* ss_ptrs, const int64_t* ss_inds, int64_t n_elim_ranges,
                          const int64_t* elim_ranges, int64_t n_elim_last, const int64_t* elim_last_ids,
                          bspb200_solver** out);

/* replaces the raw Solver constructor (reference Solver.h:37-38) over a skeleton built from
 * (spanStart, lumpToSpan, colPtr, rowInd) (reference CoalescedBlockMatrix.h:39-41). permutation may be NULL (identity). */
int bspb200_create_solver_from_skel(int backend, int num_threads, int64_t n_spans, const int64_t* span_start,
                                    int64_t n_lumps, const int64_t* lump_to_span, const int64_t* col_ptr,
                                    const int64_t* row_ind, int64_t n_elim_ranges, const int64_t* elim_ranges,
                                    const int64_t* permutation, bspb200_solver** out);

void bspb200_destroy_solver(bspb200_solver* s);

int64_t bspb200_solver_query(const bspb200_solver* s, int what);
/* copies min(len, cap) entries to out (out may be NULL), returns the full length, <0 on error */
int64_t bspb200_solver_array(const bspb200_solver* s, int which, int64_t* out, int64_t cap);

/* host-side helpers on HOST buffers: CoalescedBlockMatrixSkel::densify / damp (reference CoalescedBlockMatrix.cpp:124-187);
 * dense is row-major (order - spanStart[start_span])^2 */
int bspb200_densify(const bspb200_solver* s, int dtype, const void* host_data, void* host_dense, int fill_upper_half,
                    int64_t start_span);
int bspb200_damp(const bspb200_solver* s, int dtype, void* host_data, double alpha, double beta);
/* accessor()->blockOffset / diagBlockOffset on user block indices (reference Accessor.h:143-161) */
int bspb200_block_offset(const bspb200_solver* s, int64_t row_block, int64_t col_block, int64_t* offset,
                         int64_t* stride, int* flipped);

/* algorithmic work of the skeleton: factor flops, solve flops per RHS, nnz(L),
 * bytes moved by the sparse-elimination ranges, bytes of one solve (nRHS=1) */
int bspb200_work_estimate(const bspb200_solver* s, double* factor_flops, double* solve_flops_per_rhs, double* nnz_l,
                          double* elim_bytes_f64, double* elim_flops);

/* ---- numeric phase; `data`, `vec` are DEVICE pointers. stream = cudaStream_t (NULL = default stream). */
int bspb200_set_stream(bspb200_solver* s, void* stream);
int bspb200_set_fused(bspb200_solver* s, int enabled); /* 0: drive the fine-grained NumericCtx/SolveCtx ops one by one */

/* Solver::factor / factorUpTo / factorFrom (reference Solver.h:57, 77, 97): spans [start_span, end_span), end_span=-1: all */
int bspb200_factor(bspb200_solver* s, int dtype, void* data, int64_t start_span, int64_t end_span);
/* Solver::factor<std::vector<T*>> (reference Solver.cpp:459-460): host array of `batch` device pointers */
int bspb200_factor_batched(bspb200_solver* s, int dtype, void* const* data_ptrs, int batch, int64_t start_span,
                           int64_t end_span);
/* Solver::solve / solveL / solveLt (+UpTo/From) (reference Solver.h:61-108). vec: column-major order x n_rhs, ld */
int bspb200_solve(bspb200_solver* s, int dtype, int mode, const void* data, void* vec, int64_t ld, int n_rhs,
                  int64_t start_span, int64_t end_span);
int bspb200_solve_batched(bspb200_solver* s, int dtype, int mode, const void* const* data_ptrs, void* const* vec_ptrs,
                          int batch, int64_t ld, int n_rhs, int64_t start_span, int64_t end_span);
/* Solver::addMvFrom (reference Solver.h:88-90), Solver::pseudoFactorFrom (Solver.h:93-94) */
int bspb200_add_mv_from(bspb200_solver* s, int dtype, const void* data, int64_t span_index, const void* in_vec,
                        int64_t in_stride, void* out_vec, int64_t out_stride, int n_rhs, double alpha);
int bspb200_pseudo_factor_from(bspb200_solver* s, int dtype, void* data, int64_t span_index);
/* test hook used by the reference's own tests (CudaFactorTest.cpp:156-165):
 * createNumericCtx(0) -> doElimination(internalGetElimCtx(range_index), data, range) */
int bspb200_do_elimination(bspb200_solver* s, int dtype, void* data, int range_index);

/* Solver::deviceAccessor() (reference Solver.h:48, MatOpsCuda.cu:87-92, Accessor.h:110-200): the trivially copyable
 * accessor over DEVICE copies of the index arrays, meant to be passed BY VALUE to user kernels that write their blocks
 * straight into the factor buffer. out_ptrs receives its 8 device pointers in the member order of
 * PermutedCoalescedAccessor: spanStart, spanToLump, lumpStart, spanOffsetInLump, chainColPtr, chainRowSpan, chainData,
 * permutation (rebuild the struct with PermutedCoalescedAccessor::init; they stay valid while the solver lives). */
int bspb200_device_accessor(const bspb200_solver* s, const int64_t** out_ptrs);

/* ---- end-to-end convenience on HOST buffers (pinned or pageable): copies A up, factors, solves n_rhs
 * right-hand sides in place, copies L (if host_factor_out != NULL) and x back. The strictly-upper triangles of wide
 * diagonal blocks (a don't-care region of the format) are not uploaded; host_factor_out holds zeros or factor
 * by-products there. When the first elimination range is large its columns go up in chunks and the elimination of a
 * chunk overlaps the upload of the next (BSPB200_HOST_CHUNKS, default 12; the chunked summation order is fixed but
 * differs from factor()'s in the last bits). */
int bspb200_factor_solve_host(bspb200_solver* s, int dtype, const void* host_data, void* host_factor_out, void* host_vec,
                              int64_t ld, int n_rhs);

/* bytes the last *_host call of this solver moved in each direction (what bench.py declares as e2e traffic) */
int bspb200_host_copy_bytes(const bspb200_solver* s, int64_t* h2d_bytes, int64_t* d2h_bytes);
/* the batched form (reference Solver::factor / solve on std::vector<T*>, Solver.cpp:459-536) on HOST buffers: `batch`
 * identically structured matrices and their right-hand sides; uploads of later sub-batches overlap the factorization
 * of earlier ones; solutions are written back in place */
int bspb200_factor_solve_host_batched(bspb200_solver* s, int dtype, const void* const* host_datas, int batch,
                                      void* const* host_vecs, int64_t ld, int n_rhs);

/* ---- dense building blocks on DEVICE pointers (row-major), the kernels behind NumericCtx::potrf / trsm / saveSyrkGemm
 * (reference MatOps.h:124-130); exposed for kernel-level parity tests and roofline measurements.
 * C(m x n) = alpha * A(m x k) * B(n x k)^T + beta * C ; lower_only: only entries with col <= row are written */
int bspb200_dev_gemm_nt(int dtype, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                        const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int lower_only, void* stream);
/* in-place Cholesky of the (n + rows_below) x n trapezoid (top n x n = diagonal block, lower triangle), ld >= n */
int bspb200_dev_potrf(int dtype, int64_t n, int64_t rows_below, void* A, int64_t ld, void* stream);

/* host only: the ticket-ordered job list of the tile-DAG Cholesky kernel for a lump of block_cols x block_rows tiles of 96
 * (segment_len 0 = whole sums; see LumpCholKernel.cu buildJobs): 5 int32 per job {block row, block column, first K block,
 * end K block, type | last << 4}; returns the number of jobs (copies at most cap_jobs). No device is touched. */
int64_t bspb200_lumpchol_job_list(int block_cols, int block_rows, int segment_len, int lag, int32_t* jobs_out, int64_t cap_jobs);

/* per-kernel-class profiling (CUDA events around every launch, algorithmic flops/bytes beside the time); the
 * report is a JSON object {class: {launches, ms, flops, bytes}}; returns its length (copied up to cap-1 + NUL) */
int bspb200_profile_enable(int on);
int64_t bspb200_profile_report(char* json_out, int64_t cap);

/* diagnostics: what == 0 -> up to 64 clock64() phase stamps of CTA 0 of the last panel-kernel launch (only recorded
 * when the environment variable BSPB200_PANEL_CLK=1 is set at first use); returns the bytes copied, < 0 on error */
int64_t bspb200_debug_read(int what, void* out, int64_t bytes);

/* number of kernel launches issued by this library since process start (bench.py "gpu_launches") */
int64_t bspb200_launch_count(void);

/* ---- synthetic problems (reference baspacho/testing/TestingMatGen.cpp, TestingUtils.cpp, Bench.cpp:279-357) */
/* kind: 0 flat(size,fill) 1 grid(w,h,fill,conn) 2 meridians(num,len,fill,band,hairLen,nHairs,sHairs)
 *       3 bundle-adjustment(numPts,numCams,minObs,meanExtra,window,farProb) 4 randomCols(size,fill)
 *       5 flat+schur(size,fill,schurSize,schurFill).  Block sizes: uniform in [bsize_min,bsize_max] (points/cams for 3). */
int bspb200_gen_pattern(int kind, const double* params, int n_params, int64_t bsize_min, int64_t bsize_max, int64_t seed,
                        bspb200_pattern** out);
int64_t bspb200_pattern_order(const bspb200_pattern* p);
int64_t bspb200_pattern_nnz(const bspb200_pattern* p);
int bspb200_pattern_copy(const bspb200_pattern* p, int64_t* param_sizes, int64_t* ss_ptrs, int64_t* ss_inds);
void bspb200_pattern_free(bspb200_pattern* p);
/* randomData(size, low, high, seed) (reference TestingUtils.cpp:39-52), written to a HOST buffer */
int bspb200_random_data(int dtype, int64_t size, double low, double high, int64_t seed, void* host_out);
/* fillReducingPermutation of a CSR/CSC pattern (reference SparseStructure.cpp:313-330) */
int bspb200_fill_reducing_permutation(int64_t n, const int64_t* ptrs, const int64_t* inds, int64_t* perm_out);

#ifdef __cplusplus
}
#endif
#endif /* BASPACHO_B200_H_ */
