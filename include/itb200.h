/*
 * itb200.h — C ABI of the B200-native block-sparse contraction path (libitb200.so).
 *
 * This is the thin `extern "C"` layer that the storage-type plugin (itensor_b200/plugin/gpu_storage.h:
 * DenseGPU<T>, QDenseGPU<T> and their doTask overloads) calls. Plain pointers and sizes only.
 * All Index / QN / TagSet bookkeeping stays in the caller; what crosses this boundary is
 *   - integer block structure (sector sizes, block coordinates, element offsets), and
 *   - flat device (or host, for the *_host entry points) data pointers.
 *
 * Reference interfaces replaced (paths relative to the ITensor v3 tree):
 *   itb_contract_*      doTask(Contract&,QDense<VA>,QDense<VB>,ManageStore&)   itensor/itdata/qdense.cc:671-747
 *                       doTask(Contract&,Dense<T1>,Dense<T2>,ManageStore&)     itensor/itdata/dense.cc:262-330
 *                       getContractedOffsets / loopContractedBlocks            itensor/itdata/qutil.h:93-371
 *                       computeLabels / contractIS                             itensor/tensor/contract.h:155-202,
 *                                                                              itensor/indexset_impl.h:77-129
 *                       contract(TenRefc,Labels,TenRefc,Labels,TenRef,Labels)  itensor/tensor/contract.cc:843-866
 *                       gemm                                                   itensor/tensor/gemm.cc:253-288
 *   itb_permute_*       doTask(Order const&,QDense<T>&) / permuteQDense         itensor/itdata/qdense.cc:847-893
 *                       doTask(Order const&,Dense<T>&)  / permuteDense          itensor/itdata/dense.cc:417-439
 *                       add(PlusEQ,QDense,QDense) (permuting accumulate)        itensor/itdata/qdense.cc:515-549
 *                       transform()                                             itensor/tensor/ten_impl.h:107-160
 *   itb_nrm2            doTask(NormNoScale,QDense/Dense) -> dnrm2               qdense.cc:409-417, dense.cc:150-162
 *   itb_axpy            PlusEQ trivial-permutation fast path -> daxpy           qdense.cc:523-529, dense.cc:380-385
 *   itb_scal            doTask(Mult<Real/Cplx>,...)                             qdense.cc:303-325, dense.cc:112-136
 *   itb_fill            doTask(Fill<T>,...)                                     qdense.cc:356-373, dense.cc:90-109
 *   itb_conj            doTask(Conj,...)                                        qdense.cc:376-380, dense.cc:167-171
 *   itb_get_elt         doTask(GetElt,...) (host read-back of one element)      qdense.cc:232-245, dense.cc:37-47
 *   itb_flux_blocks     getBlockOffsets(IndexSet,QN)                            qdense.cc:133-173
 *
 * Conventions
 *   - Everything is column-major inside a block: the FIRST index is the fastest
 *     (itensor/tensor/range.h:197-206).
 *   - Block lists are sorted by the reference's Block ordering: reverse-lexicographic,
 *     i.e. the LAST index is most significant (itensor/itdata/qdense.cc:60-65).
 *   - Complex data is interleaved (re,im) doubles, exactly std::complex<double>.
 *   - A dense tensor is the special case of one sector per index and a single block.
 *   - Every function returns ITB_OK (0) or a negative error code; itb_last_error() gives the
 *     message. There is NO CPU fallback: entry points that need the device fail with
 *     ITB_ERR_CUDA when no CUDA device is usable.
 */
#ifndef ITB200_H
#define ITB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ITB_OK 0
#define ITB_ERR_INVALID (-1)     /* malformed arguments / structure */
#define ITB_ERR_CUDA (-2)        /* CUDA runtime failure (message has the CUDA string) */
#define ITB_ERR_UNSUPPORTED (-3) /* valid input outside what the kernels support */
#define ITB_ERR_NOMEM (-4)

#define ITB_MAX_ORDER 12 /* indices per tensor (reference Block = InfArray<long,11> + heap) */

enum { ITB_F64 = 0, ITB_C64 = 1 };

typedef struct itb_ctx itb_ctx;                     /* one per (process, GPU): stream + memory pool */
typedef struct itb_contract_plan itb_contract_plan; /* host-computed block-pair tables */
typedef struct itb_permute_plan itb_permute_plan;   /* host-computed block-permute tables */

/* Integer structure of a (block-sparse or dense) tensor; pointers are borrowed for the call. */
typedef struct itb_tensor_desc {
    int32_t order;          /* number of indices */
    int32_t dtype;          /* ITB_F64 | ITB_C64 */
    const int32_t* nsect;   /* [order] sectors (QN blocks) per index */
    const int64_t* sect;    /* concatenated sector sizes, sum(nsect) entries (Index::blocksize0) */
    int64_t nblocks;        /* stored blocks */
    const int32_t* blocks;  /* [nblocks*order] block coordinates, reference-sorted */
    const int64_t* offsets; /* [nblocks] element offset of each block in the flat store */
    int64_t nelems;         /* elements in the flat store (QDense::store.size()) */
} itb_tensor_desc;

typedef struct itb_contract_info {
    int32_t c_order;
    int32_t c_dtype;
    int64_t c_nblocks;
    int64_t c_nelems;
    int64_t npairs;       /* (A block, B block) pairs == reference blockContractions.size() */
    double flops;         /* sum_pairs 2*M*N*K (x2 real*cplx, x4 cplx*cplx)  — SURVEY §8(d) */
    int64_t n_gemm_tiles; /* work items of the DMMA tile kernel */
    int64_t n_skinny;     /* work items of the streaming (min(M,N) small) kernel */
    int64_t n_dot;        /* work items of the split-K reduction kernel */
    int64_t table_bytes;  /* bytes of compact tables uploaded to the device */
    double class_flops[5]; /* flops by kernel class: tile 128x128, tile 64x64, tile 32x32, streaming, split-K */
} itb_contract_info;

/* ---- library / context ------------------------------------------------------------------ */
const char* itb_version(void);
const char* itb_last_error(void);
int itb_device_count(void);
int itb_ctx_create(int device, itb_ctx** out);
int itb_ctx_destroy(itb_ctx* ctx);
void* itb_ctx_stream(itb_ctx* ctx);                 /* the cudaStream_t every launch goes to */
int itb_ctx_device(itb_ctx* ctx);                   /* CUDA device ordinal of the context */
int itb_ctx_set_stream(itb_ctx* ctx, void* stream); /* adopt a caller-owned stream */
int itb_synchronize(itb_ctx* ctx);
int64_t itb_launch_count(itb_ctx* ctx);             /* kernels launched through this context */

/* ---- device memory (stream-ordered caching pool) ---------------------------------------- */
int itb_malloc(itb_ctx* ctx, size_t bytes, void** dptr);
int itb_free(itb_ctx* ctx, void* dptr);
int itb_memcpy_h2d(itb_ctx* ctx, void* dst, const void* src, size_t bytes);
int itb_memcpy_d2h(itb_ctx* ctx, void* dst, const void* src, size_t bytes); /* synchronises */
int itb_memcpy_d2d(itb_ctx* ctx, void* dst, const void* src, size_t bytes);
int itb_memset0(itb_ctx* ctx, void* dst, size_t bytes);
int itb_pool_trim(itb_ctx* ctx);

/* ---- host-only integer planning (no device needed) --------------------------------------- */
/* labels: arbitrary ints; a label present in both labA and labB marks a contracted index. */
int itb_contract_plan_create(const itb_tensor_desc* A, const int32_t* labA,
                             const itb_tensor_desc* B, const int32_t* labB,
                             itb_contract_plan** out);
int itb_contract_plan_destroy(itb_contract_plan* plan);
int itb_contract_plan_info(const itb_contract_plan* plan, itb_contract_info* out);
/* result structure and flop count only (never builds the device tables; any out pointer may be NULL) */
int itb_contract_plan_shape(const itb_contract_plan* plan, int32_t* c_order, int32_t* c_dtype, int64_t* c_nblocks, int64_t* c_nelems,
                            int64_t* npairs, double* flops);
/* result structure, in the reference's order (contractIS with sortResult=false) */
int itb_contract_plan_c_labels(const itb_contract_plan* plan, int32_t* labels /*[c_order]*/);
int itb_contract_plan_c_nsect(const itb_contract_plan* plan, int32_t* nsect /*[c_order]*/);
int itb_contract_plan_c_sect(const itb_contract_plan* plan, int64_t* sect /*[sum nsect]*/);
int itb_contract_plan_c_blocks(const itb_contract_plan* plan, int32_t* blocks /*[c_nblocks*c_order]*/);
int itb_contract_plan_c_offsets(const itb_contract_plan* plan, int64_t* offsets /*[c_nblocks]*/);
/* (iA,iB,iC) positions into the three block lists, in the reference's enumeration order */
int itb_contract_plan_pairs(const itb_contract_plan* plan, int64_t* triples /*[npairs*3]*/);
/* restrict execution to C blocks [first,last) — the output-block sharding unit for multi-GPU */
int itb_contract_plan_set_cblock_range(itb_contract_plan* plan, int64_t first, int64_t last);
/* general form: mask[c] != 0 selects C block c (c_nblocks entries); NULL selects all. Unselected C blocks
 * are neither computed nor written (their storage is left untouched). */
int itb_contract_plan_set_cblock_mask(itb_contract_plan* plan, const uint8_t* mask);
/* finer sharding unit: restrict execution to a row-slice of ONE index of C. C blocks whose coordinate on index c_index
 * (position in the result's index list) is sector s compute and write only the sub-box lo[s] <= i < hi[s] of that index
 * (lo/hi: nsect(c_index) entries; hi[s] <= lo[s] skips those blocks; NULL/NULL removes the restriction). Everything outside
 * the sub-boxes is left untouched. This is the "rows of the primed link inside a QN sector" partition of SURVEY 8(e)
 * (reference rule: one owner per C block, itensor/itdata/qutil.h:285-348, refined to row ranges for balance). Supported when,
 * in every executed block, the sliced index is the slowest non-unit uncontracted index of the operand it comes from
 * (always true for the H_eff*phi chain with QN site indices); otherwise ITB_ERR_UNSUPPORTED and the plan is unchanged. */
int itb_contract_plan_set_index_slices(itb_contract_plan* plan, int32_t c_index, const int64_t* lo, const int64_t* hi);
/* introspection of the device work lists (tests, schedule analysis). Tile items of the DMMA kernel in queue
 * order, 8 int32 each: {C block (position in the plan's executed C-block list), m0, n0, tile_m, tile_n,
 * chunk_begin, chunk_end, ws_slot}; executed C blocks, 4 int64 each: {M, N, ksum, npairs} (real-expanded dims).
 * All return the item count and write at most cap items (out may be NULL). */
int64_t itb_contract_plan_tiles(const itb_contract_plan* plan, int32_t* out, int64_t cap);
int64_t itb_contract_plan_cblks(const itb_contract_plan* plan, int64_t* out, int64_t cap);
/* row groups of the streaming class, 4 int64 each: {input slots, output slots, long-side dims, rows} */
int64_t itb_contract_plan_rowgroups(const itb_contract_plan* plan, int64_t* out, int64_t cap);
/* stream-K partition: CTA b of the persistent grid owns tile items [cta_begin[b], cta_begin[b+1]) */
int64_t itb_contract_plan_cta_begin(const itb_contract_plan* plan, int32_t* out, int64_t cap);

/* blocks allowed by a total flux: sum_j dir[j]*qn(index j, sector) == flux (component-wise,
 * component c taken modulo |mod[c]| when |mod[c]|>1). qn: for index j, sector s, component c:
 * qn[(sect_start[j]+s)*nqn + c]. Returns the count; writes at most cap blocks (may be NULL). */
int64_t itb_flux_blocks(int32_t order, const int32_t* nsect, const int32_t* qn, int32_t nqn,
                        const int32_t* mod, const int32_t* dir, const int32_t* flux,
                        int32_t* blocks, int64_t cap);

/* ---- contraction -------------------------------------------------------------------------- */
/* C = A*B over all matching block pairs; dA,dB,dC device pointers; C is fully overwritten. */
int itb_contract_run(itb_ctx* ctx, itb_contract_plan* plan, const void* dA, const void* dB, void* dC);
/* same through HOST buffers: H2D(A,B) + run + D2H(C) (the end-to-end path) */
int itb_contract_host(itb_ctx* ctx, itb_contract_plan* plan, const void* hA, const void* hB, void* hC);

/* ---- permute / permuting accumulate ------------------------------------------------------- */
/* perm[i] = position in dst of src index i (reference Permutation::dest(i)). Every src block
 * must have its image in dst; dst blocks without a source are zero-filled when !accumulate.
 * dst dtype may be C64 with src F64 (promotion). */
int itb_permute_plan_create(const itb_tensor_desc* src, const itb_tensor_desc* dst,
                            const int32_t* perm, itb_permute_plan** out);
int itb_permute_plan_destroy(itb_permute_plan* plan);
int64_t itb_permute_plan_bytes(const itb_permute_plan* plan); /* algorithmic bytes: read src + write dst */
/* dst = alpha*P(src)   (accumulate==0)   or   dst += alpha*P(src)   (accumulate!=0) */
int itb_permute_run(itb_ctx* ctx, itb_permute_plan* plan, const void* dSrc, void* dDst,
                    double alpha_re, double alpha_im, int accumulate);
int itb_permute_host(itb_ctx* ctx, itb_permute_plan* plan, const void* hSrc, void* hDst,
                     double alpha_re, double alpha_im, int accumulate);

/* Generic strided block copies (the kernel behind QCombiner combine/uncombine, itensor/itdata/qcombiner.cc:125-301,
 * and any sub-block scatter/gather): item i copies an n-dim box  dst[d_off + sum_j i_j*dstr_j] = src[s_off + sum_j i_j*sstr_j].
 * Offsets/strides in ELEMENTS of the respective dtype. The returned plan is run with itb_permute_run (alpha, accumulate
 * as there; nothing is zero-filled — use itb_memset0 on the destination first if it needs zeros). */
typedef struct itb_copy_item {
    int64_t s_off, d_off;
    int32_t n, pad_;
    int64_t ext[ITB_MAX_ORDER];
    int64_t sstr[ITB_MAX_ORDER];
    int64_t dstr[ITB_MAX_ORDER];
} itb_copy_item;
int itb_blockcopy_plan_create(int64_t nitems, const itb_copy_item* items, int32_t src_dtype, int32_t dst_dtype,
                              itb_permute_plan** out);

/* ---- BLAS-1 style storage tasks (n counts ELEMENTS of the given dtype) -------------------- */
int itb_nrm2(itb_ctx* ctx, int32_t dtype, int64_t n, const void* dX, double* out); /* syncs */
int itb_scal(itb_ctx* ctx, int32_t dtype, int64_t n, void* dX, double alpha_re, double alpha_im);
int itb_axpy(itb_ctx* ctx, int32_t dtype, int64_t n, double alpha_re, double alpha_im,
             const void* dX, void* dY); /* Y += alpha X, same dtype */
int itb_fill(itb_ctx* ctx, int32_t dtype, int64_t n, void* dX, double re, double im);
int itb_conj(itb_ctx* ctx, int64_t n, void* dX);                 /* C64 in place */
int itb_real_to_cplx(itb_ctx* ctx, int64_t n, const void* dX, void* dY);
int itb_take_part(itb_ctx* ctx, int64_t n, const void* dX, void* dY, int imag);
int itb_get_elt(itb_ctx* ctx, int32_t dtype, const void* dX, int64_t offset, double out[2]); /* syncs */
int itb_dot(itb_ctx* ctx, int32_t dtype, int64_t n, const void* dX, const void* dY, int conj_x,
            double out[2]); /* syncs */

/* ---- per-block eigh / SVD for the svdBond path, host buffers in/out ------------------------ */
/* LAPACK dsyev('V','U') / zheev semantics (itensor/tensor/lapack_wrap.cc:322-347,678-706): A (n x n, column-major)
 * is overwritten by the eigenvectors, w receives the eigenvalues in ascending order. cuSOLVER syevd. */
int itb_syevd_host(itb_ctx* ctx, int32_t dtype, int32_t n, void* hA, double* hW, int32_t* info);
/* LAPACK gesdd(jobz='S') semantics (lapack_wrap.cc:365-486): thin SVD A = U diag(s) VT of an m x n column-major
 * matrix (destroyed); U is m x l (ldu=m), VT is l x n (ldvt=l), l=min(m,n). cuSOLVER gesvd. */
int itb_gesvd_host(itb_ctx* ctx, int32_t dtype, int32_t m, int32_t n, void* hA, double* hS, void* hU, void* hVT, int32_t* info);

/* 1 once the device solvers can be used without first-touch stalls. cuSOLVER / cuBLAS load their kernels lazily; on a
 * box whose page cache is cold the context therefore reads those shared objects sequentially in a background thread
 * (started by itb_ctx_create, ITB_WARM_LIBS=0 disables). Callers keep small-matrix work on host LAPACK until this
 * returns 1 (plugin/lapack_gpu.cc, plugin/svd_gpu.cc). */
int itb_solver_ready(void);

/* Device-resident batched SVD of the blocks of an order-2 block-sparse tensor (svdOrd2 / svdImpl QN loop,
 * itensor/svd.cc:169-422, on QDenseGPU storage): block b is an m[b] x n[b] column-major matrix at ELEMENT offset
 * a_off[b] of the device buffer dA (not modified). A_b = U_b diag(s_b) V_b^H with l_b = min(m,n) singular values in
 * descending order. U and V stay on the device inside the batch object; the caller reads the singular values back
 * (itb_svd_batch_values: concatenated in block order, synchronises), decides the truncation on the host and copies
 * the kept leading columns into the new tensors with itb_svd_batch_copy_u / _v (device to device, ordered on the
 * context's stream). cuSOLVER Xgesvdp (polar decomposition) for min(m,n) >= 96, gesvdj below, 4 concurrent streams. */
typedef struct itb_svd_batch itb_svd_batch;
int itb_svd_batch_run(itb_ctx* ctx, int32_t dtype, int64_t nblocks, const int64_t* a_off, const int32_t* m, const int32_t* n,
                      const void* dA, itb_svd_batch** out);
int itb_svd_batch_values(itb_svd_batch* batch, double* hS);
int itb_svd_batch_copy_u(itb_svd_batch* batch, int64_t block, int32_t ncols, void* dDst);           /* m x ncols */
int itb_svd_batch_copy_v(itb_svd_batch* batch, int64_t block, int32_t ncols, void* dDst, int conj); /* n x ncols */
int itb_svd_batch_destroy(itb_svd_batch* batch);
/* diagnostics since process start: out = {blocks given to the polar solver, of those redone with Jacobi (failed or perturbed
 * input, err_sigma > 1e-14 s0), blocks factorised with Jacobi}; returns the largest err_sigma / s0 the polar solver reported */
double itb_svd_batch_stats(int64_t out[3]);

/* Device-resident batched eigh of the (square, Hermitian) blocks of an order-2 block-sparse tensor — the per-block loop of
 * diagHImpl (itensor/hermitian.cc:231-257) on QDenseGPU storage, i.e. the density-matrix branch of svdBond (noise > 0).
 * Block b is n[b] x n[b] at ELEMENT offset a_off[b] of dA (not modified). negate != 0 diagonalises -A_b, so that the ascending
 * order of syevd is the reference's largest-first order (tensor/algs_impl.h:123-137; the caller flips the sign of the values).
 * Eigenvectors stay on the device; itb_eigh_batch_values reads the eigenvalues back (block order, synchronises), the caller
 * truncates on the host and copies the kept leading eigenvectors with itb_eigh_batch_copy_vectors (device to device). */
typedef struct itb_eigh_batch itb_eigh_batch;
int itb_eigh_batch_run(itb_ctx* ctx, int32_t dtype, int64_t nblocks, const int64_t* a_off, const int32_t* n, const void* dA, int negate,
                       itb_eigh_batch** out);
int itb_eigh_batch_values(itb_eigh_batch* batch, double* hW);
int itb_eigh_batch_copy_vectors(itb_eigh_batch* batch, int64_t block, int32_t ncols, void* dDst, int conj); /* n x ncols */
int itb_eigh_batch_destroy(itb_eigh_batch* batch);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch ---------------------------------------------
 * The path shards by ROWS of one uncontracted index (itb_contract_plan_set_index_slices); the only exchange is re-replicating
 * a row-sharded tensor: pack own rows (itb_blockcopy_plan_create) -> itb_comm_allgather -> scatter the other ranks' rows.
 * NCCL (libnccl.so.2) is loaded on first use. Rank 0 obtains an id with itb_comm_unique_id and hands it to the other processes
 * (the plugin does that through a file, plugin/gpu_storage.cc); every process then calls itb_comm_create (collective). */
#define ITB_COMM_ID_BYTES 128
typedef struct itb_comm itb_comm;
int itb_comm_unique_id(uint8_t out[ITB_COMM_ID_BYTES]);
int itb_comm_create(itb_ctx* ctx, int32_t world, int32_t rank, const uint8_t id[ITB_COMM_ID_BYTES], itb_comm** out);
int32_t itb_comm_world(const itb_comm* comm);
int32_t itb_comm_rank(const itb_comm* comm);
/* every rank contributes count doubles at dSend; dRecv receives world*count doubles in rank order (in place when
 * dSend == (double*)dRecv + rank*count); ordered on the context's stream */
int itb_comm_allgather(itb_comm* comm, itb_ctx* ctx, const void* dSend, void* dRecv, int64_t count);
int itb_comm_destroy(itb_comm* comm);
/* The contraction with its result additionally stored into n <= 7 peer copies of C (other ranks' buffers mapped with
 * itb_p2p_open; same layout as dC): the row exchange of a multi-GPU product rides the epilogue of the kernel that produces
 * the rows, tile by tile over NVLink, instead of following it as a collective (the reference's OpenMP workers write into
 * one shared C, itensor/itdata/qutil.h:285-348). ITB_ERR_UNSUPPORTED when the plan has work outside the static DMMA tile class
 * (nothing has been launched then): push the rows with a block-copy plan after itb_contract_run instead. */
int itb_contract_run_mirrored(itb_ctx* ctx, itb_contract_plan* plan, const void* dA, const void* dB, void* dC, int32_t n, void* const* peer_dC);
/* Peer memory for the direct row exchange (no counterpart in the reference: its OpenMP workers share one address space,
 * itensor/itdata/qutil.h:285-348). itb_p2p_alloc: a device buffer other ranks may map + its 64-byte export handle (publish it
 * through any host channel); itb_p2p_open: map a peer's buffer into this process (enables peer access over NVLink from the
 * context's device); the mapped pointer is a plain device pointer for every itb_* call, e.g. as the destination of a
 * block-copy plan whose items carry the rows this rank owns. ITB_ERR_UNSUPPORTED where CUDA IPC / peer access is missing
 * (callers fall back to itb_comm_allgather). */
#define ITB_P2P_HANDLE_BYTES 64
int itb_p2p_alloc(itb_ctx* ctx, int64_t bytes, void** dptr, uint8_t handle[ITB_P2P_HANDLE_BYTES]);
int itb_p2p_open(itb_ctx* ctx, const uint8_t handle[ITB_P2P_HANDLE_BYTES], void** peer_ptr);
int itb_p2p_close(itb_ctx* ctx, void* peer_ptr);
int itb_p2p_free(itb_ctx* ctx, void* dptr);
/* arrival barrier over peer memory: a flag block of ITB_P2P_FLAG_BYTES per rank (itb_p2p_alloc + itb_p2p_barrier_init),
 * mapped by every peer; one tiny launch per barrier on the context's stream (CUDA-graph friendly: the epoch lives in the
 * block). Everything this rank stored into peer buffers before the call is visible to a peer once its barrier returns.
 * peer_flags[r] = rank r's block as mapped here (entry `rank` ignored). The wait is bounded; itb_p2p_barrier_status
 * reports an expired wait. */
#define ITB_P2P_MAX_WORLD 8
#define ITB_P2P_FLAG_BYTES 256
int itb_p2p_barrier_init(itb_ctx* ctx, void* local_flags);
int itb_p2p_barrier(itb_ctx* ctx, void* local_flags, void* const* peer_flags, int32_t world, int32_t rank);
int itb_p2p_barrier_status(itb_ctx* ctx, const void* local_flags, int64_t* epoch, int64_t* error);
/* modelled device work of a plan as currently restricted (row slices included): cycles of the DMMA tile class summed over all
 * CTAs of the persistent grid, and algorithmic bytes of the HBM-bound streaming class — what a multi-GPU row partition
 * balances (a cut that leaves a short remainder tile costs almost a full tile per K-chunk; flops do not see that) */
int itb_contract_plan_model_work(const itb_contract_plan* plan, double* tile_cycles, double* stream_bytes);
/* flops of every C block of a plan (2*M*N*K summed over its pairs, complex multipliers included): the weights of a row partition */
int itb_contract_plan_cblock_flops(const itb_contract_plan* plan, double* out /*[c_nblocks]*/);

/* Measured refinement of a plan's static tile partition (no counterpart in the reference: its OpenMP loop over C blocks,
 * itensor/itdata/qutil.h:285-348, is scheduled dynamically by the runtime). Executes the plan rounds+1 times on the given
 * operands (dC holds the correct result afterwards), reads the per-CTA clock64 spans of the tile kernel after each run,
 * rescales the modelled cost of every tile by measured/modelled and re-partitions; keeps the partition with the shortest
 * longest span. *gain (optional) = longest span before / after. Synchronises the stream; meant for plans that are executed
 * many times (Davidson repeats the same structures). Plans without tile items are left as they are. */
int itb_contract_plan_refine(itb_ctx* ctx, itb_contract_plan* plan, const void* dA, const void* dB, void* dC, int rounds, double* gain);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* profile!=0: itb_contract_run brackets every kernel launch with CUDA events (adds syncs; measurement
 * only). itb_contract_last_ms then returns the device time of the last run by kernel class
 * (same order as itb_contract_info.class_flops; 0 for classes that did not launch). */
int itb_ctx_set_profile(itb_ctx* ctx, int profile);
int itb_contract_last_ms(itb_ctx* ctx, float ms[5]);
/* profile mode: clock64 span of every CTA of the last DMMA tile-kernel launch (schedule calibration); returns the count */
int64_t itb_contract_last_cta_cycles(itb_ctx* ctx, int64_t* out, int64_t cap);
/* profile mode: per queue item of the last DMMA tile-kernel launch, 4 int64 each: {CTA that ran it, clock64 at item start, at the
 * end of its K loop, at the end of its epilogue} (relative to the CTA's start); returns the item count, writes at most cap int64 */
int64_t itb_contract_last_item_cycles(itb_ctx* ctx, int64_t* out, int64_t cap);
/* bare DMMA (mma.sync f64) and DFMA issue loops, no memory traffic; returns TFLOP/s */
int itb_peak_fp64(itb_ctx* ctx, int which /*0=dmma m8n8k4, 1=dfma, 2=dmma m16n8k8*/, int iters, double* tflops);
/* device-side timing on the context's stream (CUDA events) */
int itb_timer_start(itb_ctx* ctx);
int itb_timer_stop_ms(itb_ctx* ctx, float* ms); /* syncs */

#ifdef __cplusplus
}
#endif
#endif /* ITB200_H */
