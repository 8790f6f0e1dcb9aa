/*
 * rimu_b200.h -- C ABI of the B200-native FCIQMC step (librimu_b200.so).
 *
 * This is the drop-in boundary for Rimu.jl's hot path.  Every entry point names the
 * reference interface it replaces (path:line under RimuQMC/Rimu.jl v0.14.0 `src/`).
 * The Julia-side binding (`ccall`) a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C: opaque handles, plain pointers and sizes; no C++/torch types.
 *   - every function returns an int status: 0 ok; >0 recoverable (caller may regrow and
 *     retry, the source vector is never modified by a failed step); <0 fatal.
 *     rimu_last_error() returns a human-readable message for the calling thread.
 *   - host pointers are borrowed for the duration of the call only.
 *   - one host thread drives one context; a context owns one GPU, one stream, one working
 *     table ("working memory", the analogue of PDWorkingMemory) and optionally one NCCL
 *     communicator.  There is NO CPU fallback: without a CUDA device every compute call fails.
 *
 * Address ("key") interchange format: W little-endian uint64 words per address, word 0 least
 * significant.  BoseFS{N,M}: bit layout of BitStringAddresses/bitstring.jl:464-472 (mode 1 in
 * the lowest bits, n ones then a 0 separator), B = N+M-1 bits, W = ceil((B+1)/64) <= 2 (one spare
 * bit is kept so that the all-ones word pattern can mark empty table slots).
 * FermiFS{N,M}: bit m-1 <-> mode m (bitstring.jl:713-723).  CompositeFS of two FermiFS
 * (FermiFS2C): component c occupies bits [c*M, (c+1)*M), 2M <= 64.
 * General CompositeFS (multicomponent.jl:10-19; RIMU_ADDR_COMPOSITE, HubbardRealSpace): up to RIMU_MAX_COMPONENTS
 * BoseFS / FermiFS components over the same M modes, packed side by side from the low bits: component c occupies
 * bits [off_c, off_c + b_c) with b_c = N_c + M - 1 (BoseFS layout) or M (FermiFS layout) and off_c = b_0 + ... + b_{c-1};
 * at most 127 bits in total, W = ceil((bits + 1) / 64).  (Julia keeps one BitString per component; the shim concatenates.)
 * Julia's BitString stores chunks most-significant first (bitstring.jl:72-75); the shim
 * reverses chunk order and widens sub-64-bit chunk types.
 * Values are one 8-byte lane: double (RIMU_VAL_F64) or int64 (RIMU_VAL_I64).
 */
#ifndef RIMU_B200_H
#define RIMU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIMU_MAX_MODES 128
#define RIMU_MAX_TABLE_MODES 64
#define RIMU_MAX_COMPONENTS 4

/* status codes */
enum {
    RIMU_OK = 0,
    RIMU_ERR_TABLE_FULL = 1,     /* working table too small for this step: grow / raise active slots, retry */
    RIMU_ERR_VECTOR_FULL = 2,    /* destination vector capacity too small; *needed written where documented */
    RIMU_ERR_EXCHANGE_FULL = 3,  /* per-peer spawn exchange buffer too small */
    RIMU_ERR_WORKMEM = 4,        /* working memory of the partitioned step (bucket record streams) cannot be grown any further:
                                  * NOT helped by rimu_ctx_resize_table -- free device memory or use fewer walkers per GPU */
    RIMU_ERR_INVALID = -1,       /* bad argument / unsupported combination (ArgumentError in the reference) */
    RIMU_ERR_CUDA = -2,
    RIMU_ERR_NCCL = -3,
    RIMU_ERR_NO_DEVICE = -4
};

enum { RIMU_ADDR_BOSE = 0, RIMU_ADDR_FERMI = 1, RIMU_ADDR_FERMI2C = 2,
       RIMU_ADDR_COMPOSITE = 3 /* CompositeFS of 2..RIMU_MAX_COMPONENTS BoseFS / FermiFS components (HubbardRealSpace) */ };
enum { RIMU_HUBBARD_REAL_1D = 0, RIMU_HUBBARD_MOM_1D = 1, RIMU_HUBBARD_REAL_SPACE = 2, RIMU_TRANSCORRELATED_1D = 3,
       RIMU_HUBBARD_REAL_1D_EP = 4,        /* Hamiltonians/HubbardReal1DEP.jl:47-92: potential[] = eps_i, bosons */
       RIMU_EXTENDED_HUBBARD_REAL_1D = 5,  /* Hamiltonians/ExtendedHubbardReal1D.jl:30-135: v = neighbour interaction, bosons */
       RIMU_EXTENDED_HUBBARD_MOM_1D = 6,   /* Hamiltonians/ExtendedHubbardMom1D.jl:37-117 (bosons, boundary_condition = 0): u, v, t, kes;
                                            * ws[q] = cos(q * 2pi / M) (off-diagonals), us[d] = cos(d * (2pi / M)) (diagonal), q, d = 0..M-1 */
       RIMU_HUBBARD_MOM_1D_EP = 7          /* Hamiltonians/HubbardMom1DEP.jl:68-257 (BoseFS, two FermiFS components): u, t, kes;
                                            * potential[k] = ep[k+1], the momentum-space harmonic potential */ };
/* ExtendedHubbardReal1D boundary_condition (real ones; a complex twist angle has no device path) */
enum { RIMU_BC_PERIODIC = 0, RIMU_BC_HARD_WALL = 1, RIMU_BC_TWISTED = 2 };
enum { RIMU_VAL_F64 = 0, RIMU_VAL_I64 = 1 };
/* StochasticStyles/styles.jl: IsDeterministic (:76-105), IsStochasticInteger (:11-25),
 * IsDynamicSemistochastic (:175-214), IsStochasticWithThreshold (:117-130) */
enum { RIMU_STYLE_DETERMINISTIC = 0, RIMU_STYLE_INTEGER = 1, RIMU_STYLE_SEMISTOCHASTIC = 2, RIMU_STYLE_WITH_THRESHOLD = 3 };
/* annihilation / step methods (north_star: chosen from measured HBM GB/s):
 *   HASH      one open-addressing table in HBM, CAS claim + RED accumulate per deposit
 *   SORT      radix sort by address + segmented reduce (deterministic summation order)
 *   PARTITION spawns are appended to per-bucket record streams (bucket = fastrange of the address hash)
 *             and every bucket is annihilated in shared memory; all HBM traffic is streaming (default) */
enum { RIMU_ANNIHILATE_HASH = 0, RIMU_ANNIHILATE_SORT = 1, RIMU_ANNIHILATE_PARTITION = 2 };

typedef struct rimu_ctx rimu_ctx;
typedef struct rimu_ham rimu_ham;
typedef struct rimu_vec rimu_vec;

/* Hamiltonian descriptor.  Replaces the Julia structs
 *   HubbardReal1D (Hamiltonians/HubbardReal1D.jl:23-31), HubbardMom1D (HubbardMom1D.jl:43-65),
 *   HubbardRealSpace (HubbardRealSpace.jl:166-241), Transcorrelated1D (Transcorrelated1D.jl:55-89)
 * and the address type parameters (BoseFS{N,M}, FermiFS{N,M}, CompositeFS).
 * Tables (kes, ws, us, potential) are the SVector fields the Julia constructors precompute;
 * the host shim passes them through unchanged so both sides use identical constants. */
typedef struct {
    int32_t model;          /* RIMU_HUBBARD_* / RIMU_TRANSCORRELATED_1D */
    int32_t addr_kind;      /* RIMU_ADDR_* */
    int32_t num_modes;      /* M (per component) */
    int32_t num_components; /* 1 or 2; 2..RIMU_MAX_COMPONENTS for RIMU_ADDR_COMPOSITE */
    int32_t num_particles[2];
    int32_t ndim;           /* HubbardRealSpace: CubicGrid{D} (geometry.jl:45-54) */
    int32_t dims[3];
    int32_t fold[3];        /* periodic flag per dimension */
    int32_t cutoff;         /* Transcorrelated1D */
    int32_t three_body_term;
    int32_t has_potential;  /* potential[] valid (HubbardRealSpace trap) */
    int32_t boundary_condition; /* RIMU_BC_* (ExtendedHubbardReal1D only) */
    double u, t, v;         /* 1-D models: u,t ; Transcorrelated1D: t,v */
    double t_comp[2];       /* HubbardRealSpace hopping per component */
    double u_mat[4];        /* HubbardRealSpace interactions, column major u[i + 2*j] */
    double kes[RIMU_MAX_TABLE_MODES];
    double ws[RIMU_MAX_TABLE_MODES];
    double us[RIMU_MAX_TABLE_MODES];
    double potential[2 * RIMU_MAX_MODES]; /* potential[c*M + site] */
    /* RIMU_ADDR_COMPOSITE only (HubbardRealSpace{C} over a general CompositeFS, HubbardRealSpace.jl:18-75,166-241,340-391):
     * per-component address kind and particle number, hopping strengths t[c] and the symmetric interaction matrix
     * u[i + C*j] (column major, C = num_components); num_particles / t_comp / u_mat above are ignored for this kind */
    int32_t comp_kind[RIMU_MAX_COMPONENTS];      /* RIMU_ADDR_BOSE or RIMU_ADDR_FERMI */
    int32_t comp_particles[RIMU_MAX_COMPONENTS];
    double comp_t[RIMU_MAX_COMPONENTS];
    double comp_u[RIMU_MAX_COMPONENTS * RIMU_MAX_COMPONENTS];
} rimu_ham_desc;

/* One FCIQMC step = Interfaces.apply_operator!(wm, target, source, op, boost)
 * (Interfaces/dictvectors.jl:90-140; PDVec method DictVectors/pdworkingmemory.jl:297-309)
 * with op = FirstOrderTransitionOperator(H, shift, time_step) (fciqmc.jl:78-112), or op = H
 * itself when plain_h != 0 (mul!, DictVectors/pdvec.jl:810-822). */
typedef struct {
    int32_t style;            /* RIMU_STYLE_* */
    int32_t plain_h;
    double shift, time_step, boost;
    double proj_threshold;    /* on-the-fly projection threshold of the spawning strategy (spawning.jl:9-30) */
    double rel_threshold;     /* DynamicSemistochastic.rel_threshold (spawning.jl:358-362) */
    double abs_threshold;     /* DynamicSemistochastic.abs_threshold; +Inf = off */
    double compress_threshold;/* ThresholdCompression.threshold (compression.jl:8-26); 0 = NoCompression */
    uint64_t seed;            /* Philox key material: per-step key = f(seed, step) */
    uint64_t step;
    uint64_t table_slots;     /* working-table slots to use this step (power of two, <= ctx capacity); 0 = auto */
    /* InitiatorRule of the target vector (DictVectors/initiators.jl:132-236; PDVec(...; initiator=...) pdvec.jl:156-163):
     * deposits are kept in (safe, unsafe, initiator) lanes during annihilation -- a diagonal deposit of an initiator
     * (|parent value| > threshold) is "initiator", its spawns are "safe", spawns of non-initiators are "unsafe" -- and
     * converted back with from_initiator_value before compression.  Partitioned method only. */
    int32_t initiator_rule;   /* RIMU_INITIATOR_* ; 0 = NonInitiator */
    int32_t ordered;          /* != 0 (Float64 styles, partitioned method, no initiator rule): order-deterministic summation --
                               * every address is summed in sorted (address, value) order and the walker number in a fixed
                               * order, so that a step's result is bit-identical from run to run (audit mode, ~10x slower
                               * annihilation).  Integer walkers are always exact. */
    double initiator_threshold;
} rimu_step_params;

enum { RIMU_INITIATOR_NONE = 0, RIMU_INITIATOR = 1, RIMU_INITIATOR_SIMPLE = 2, RIMU_INITIATOR_COHERENT = 3 };

/* step_stats of the styles (styles.jl:14-20,94-96,203-209; compression.jl:16) plus
 * walkernumber_and_length of the result (pdvec.jl:896-902).  Integer-style quantities are
 * exact in the i* fields; float styles use the double fields. Sums are GLOBAL over ranks
 * when a communicator is attached. */
typedef struct {
    int64_t exact_steps, inexact_steps, spawn_attempts, len_before, len;
    double spawns, deaths, clones, zombies, norm1;
    int64_t ispawns, ideaths, iclones, izombies, inorm1;
    int64_t local_len;        /* entries stored on this rank */
    int64_t sent_records;     /* records this rank sent to peers */
    int64_t deposits;         /* non-zero deposits into the working table (diagonal + spawns), global */
    /* CUDA-event phase timings of this call: diagonal/count+scan, spawn kernel, exchange, compaction */
    float ms_diag, ms_spawn, ms_exchange, ms_compact;
    float ms_total, ms_reduce; /* ms_reduce: statistics all-reduce + read-back after the merge (multi-GPU) */
    int64_t buckets;          /* bucket count of the partitioned step on this rank (0: table method) */
    int64_t max_bucket_fill;  /* fullest bucket (parents + records) */
} rimu_step_stats;

/* ---- context ------------------------------------------------------------ */
const char *rimu_last_error(void);
int rimu_version(void);
/* struct sizes as compiled, so that FFI mirrors (Julia `struct`s, ctypes) can assert their layout */
int rimu_sizeof_ham_desc(void);
int rimu_sizeof_step_params(void);
int rimu_sizeof_step_stats(void);
/* table_slots: capacity of the working table in slots (rounded up to a power of two);
 * words: uint64 words per address (1 or 2). */
int rimu_ctx_create(int device, int words, uint64_t table_slots, rimu_ctx **out);
/* Frees the working memory.  Vectors created on the context may be destroyed afterwards (host finalizers run in any
 * order): the context's bookkeeping lives on until the last of them is gone; any other call on such a vector fails with
 * RIMU_ERR_INVALID. */
int rimu_ctx_destroy(rimu_ctx *ctx);
int rimu_ctx_synchronize(rimu_ctx *ctx);
/* Make the context's GPU the calling thread's current CUDA device (rimu_ham_create places its tables on the current
 * device; hosts that juggle several devices or libraries call this first). */
int rimu_ctx_make_current(rimu_ctx *ctx);
int rimu_ctx_table_slots(rimu_ctx *ctx, uint64_t *out);
/* reallocate the working table (contents are scratch between calls) */
int rimu_ctx_resize_table(rimu_ctx *ctx, uint64_t table_slots);
/* raw cudaStream_t of the context, for callers that want to time with their own events */
int rimu_ctx_stream(rimu_ctx *ctx, void **stream_out);
/* number of CUDA kernels this context has launched for rimu_step calls (benchmark bookkeeping) */
int rimu_ctx_launch_count(rimu_ctx *ctx, uint64_t *out);
/* step method of rimu_step: RIMU_ANNIHILATE_PARTITION (default) or RIMU_ANNIHILATE_HASH.  The environment
 * variable RIMU_B200_METHOD=hash selects the table method at context creation. */
int rimu_ctx_set_method(rimu_ctx *ctx, int method);
int rimu_ctx_get_method(rimu_ctx *ctx, int *method);
/* pinned host memory for callers that stage vectors across PCIe (cudaMallocHost / cudaFreeHost) */
int rimu_host_alloc(uint64_t bytes, void **out);
int rimu_host_free(void *p);

/* ---- multi-GPU: replaces DictVectors/communicators.jl AllToAll (:546-606),
 * mpi_exchange_alltoall! (:475-498) and merge_remote_reductions (:56) ------ */
int rimu_comm_unique_id(void *id128);               /* ncclGetUniqueId; broadcast by the host launcher */
int rimu_comm_init(rimu_ctx *ctx, const void *id128, int rank, int nranks, uint64_t exchange_records_per_peer);
int rimu_comm_rank(rimu_ctx *ctx, int *rank, int *nranks);
/* staged exchange only: after RIMU_ERR_EXCHANGE_FULL (reported identically on every rank) *needed = largest per-peer
 * record count seen; every rank then calls rimu_comm_reserve with the same larger size and repeats the step.  The direct
 * exchange sizes its streams itself. */
int rimu_comm_capacity(rimu_ctx *ctx, uint64_t *per_peer_out, uint64_t *needed_out);
int rimu_comm_reserve(rimu_ctx *ctx, uint64_t exchange_records_per_peer);
/* 1 when the spawn exchange is direct: records are bucketed by the sender and pushed straight into the owner's bucket
 * sub-streams through NVLink peer memory mapped with CUDA IPC (no receive pass); 0 = staged exchange with NCCL grouped
 * send/recv (RIMU_B200_P2P=0 forces it).  Replaces the choice between the reference's AllToAll / PointToPoint / OneSided
 * communicators (communicators.jl:107-131). */
int rimu_comm_p2p(rimu_ctx *ctx, int *enabled_out);
/* Collective teardown of the peer mappings of a multi-GPU context (call on every rank, in the same order, before
 * rimu_ctx_destroy): waits until no rank is still reading this rank's record streams, closes every imported mapping, and only
 * then lets the exporters free their buffers (CUDA requires importers to close first).  No-op on one rank; the context remains
 * usable afterwards only for local operations.  Reference: MPI.Finalize ordering, mpi_helpers.jl:9-30. */
int rimu_comm_detach(rimu_ctx *ctx);
int rimu_comm_allreduce_f64(rimu_ctx *ctx, double *host_inout, int n);
/* owner rank of an address: communicators.jl:77-81 target_segment */
int rimu_addr_owner(const uint64_t *key, int words, int nranks);
uint64_t rimu_addr_hash(const uint64_t *key, int words);

/* ---- Hamiltonian (Interfaces/hamiltonians.jl:143-201,350-370) ------------ */
int rimu_ham_create(const rimu_ham_desc *desc, rimu_ham **out);
int rimu_ham_destroy(rimu_ham *ham);
int rimu_ham_words(const rimu_ham *ham);
/* element-wise hooks evaluated by the DEVICE code, host buffers in/out:
 * diagonal_element(h, addr), num_offdiagonals(h, addr), get_offdiagonal(h, addr, chosen) for
 * chosen = first..first+count-1 (1-based as in the reference). */
int rimu_ham_diagonal(rimu_ctx *ctx, const rimu_ham *ham, const uint64_t *keys, int64_t n, double *out);
int rimu_ham_num_offdiagonals(rimu_ctx *ctx, const rimu_ham *ham, const uint64_t *keys, int64_t n, int64_t *out);
int rimu_ham_offdiagonals(rimu_ctx *ctx, const rimu_ham *ham, const uint64_t *key, int64_t first, int64_t count,
                          uint64_t *keys_out, double *vals_out);

/* ---- walker vector: replaces DVec (DictVectors/dvec.jl:44-47) / PDVec (pdvec.jl:156-163)
 * storage; dense (keys, values) arrays resident in HBM ---------------------- */
int rimu_vec_create(rimu_ctx *ctx, int val_type, uint64_t capacity, rimu_vec **out);
int rimu_vec_destroy(rimu_vec *v);
int rimu_vec_reserve(rimu_vec *v, uint64_t capacity);      /* grow, keeping contents */
int rimu_vec_clear(rimu_vec *v);                           /* zerovector!/empty! */
int rimu_vec_length(rimu_vec *v, int64_t *out);            /* length(localpart(v)) */
int rimu_vec_capacity(rimu_vec *v, uint64_t *out);
/* bucket segmentation of the partitioned step: number of buckets the vector is currently
 * segmented for (0 = none), and an explicit re-segmentation (contents unchanged; order changes) */
int rimu_vec_buckets(rimu_vec *v, uint32_t *nb_out);
int rimu_vec_rebucket(rimu_vec *v, uint32_t nb);
/* copies seg_start[nb] / seg_len[nb] to the host (test hook: every bucket's entries are contiguous) */
int rimu_vec_segments(rimu_vec *v, uint64_t *start_out, uint32_t *len_out);
/* DVec(pairs...): duplicates are summed, zeros dropped (dvec.jl:62-100); with a communicator
 * attached, keys not owned by this rank are dropped (pdvec.jl:336-349) */
int rimu_vec_upload(rimu_vec *v, const uint64_t *keys, const void *vals, int64_t n);
/* pairs(localpart(v)) in unspecified order; n_out = length */
int rimu_vec_download(rimu_vec *v, uint64_t *keys_out, void *vals_out, int64_t cap, int64_t *n_out);
int rimu_vec_copy(rimu_vec *dst, rimu_vec *src);           /* copy!/copyto! */
/* raw copyto! of n distinct non-zero pairs (e.g. a previous download); no deduplication */
int rimu_vec_assign(rimu_vec *v, const uint64_t *keys, const void *vals, int64_t n);
int rimu_vec_get(rimu_vec *v, const uint64_t *key, void *val_out); /* getindex, 0 if absent */
/* norm(v, p) p in {1, 2, inf(0)} (abstractdvec.jl:200-256); walkernumber = p=1; global with comm */
int rimu_vec_norm(rimu_vec *v, int p, double *out);
int rimu_vec_scale(rimu_vec *v, double alpha);             /* scale!/lmul! (pdvec.jl:714-729) */
/* dot(x, y) (pdvec.jl:760-796); also serves FrozenDVec dot for projected energy */
int rimu_vec_dot(rimu_vec *x, rimu_vec *y, double *out);
/* dot(::FrozenDVec, v) (pdvec.jl:773-779; freeze projectors.jl:164): n host-side (key, value) pairs against the device
 * vector -- the projected-energy reports of every step (poststepstrategy.jl:110-121).  Each key is looked up in its own
 * bucket segment only (a full scan when the vector is not segmented); global over ranks.  values are Float64. */
int rimu_vec_dot_sparse(rimu_vec *v, const uint64_t *keys, const double *values, int64_t n, double *out);
/* out = alpha*x + beta*y (add!/axpy!/axpby!, pdvec.jl:731-758); out may alias x or y */
int rimu_vec_axpby(double alpha, rimu_vec *x, double beta, rimu_vec *y, rimu_vec *out);

/* annihilation of a given spawn list (the "sum by key, drop zeros" core of
 * collect_local!/move_and_compress!, pdworkingmemory.jl:228-273).
 * method: RIMU_ANNIHILATE_HASH (open-addressing HBM table), RIMU_ANNIHILATE_SORT (radix sort + segmented reduce;
 * deterministic summation order; one-word addresses) or RIMU_ANNIHILATE_PARTITION (bucket streams + shared-memory
 * merge, what rimu_step uses; measured comparison in profiles/r1_annihilation_methods.md). Host arrays in. */
int rimu_annihilate(rimu_vec *dst, const uint64_t *keys, const void *vals, int64_t n, int method);
/* same with the spawn list already resident in HBM (device pointers), for measurement */
int rimu_annihilate_device(rimu_vec *dst, const uint64_t *d_keys, const void *d_vals, int64_t n, int method,
                           float *ms_out);

/* ---- dense-indexed deterministic H*v over a complete sector (BASELINE config 3) ----------------------------------------
 * When a Krylov vector fills a whole particle-number sector, a dictionary is the wrong container: every address is present.
 * A sector numbers all addresses of the Hamiltonian's address type (BoseFS{N,M}, FermiFS{N,M}, two FermiFS components) by
 * their combinadic rank; vectors over it are plain arrays of `dim` doubles in HBM (opaque device pointers to the host), and
 * y = H x is a gather over the off-diagonals of each address -- the same device functions the FCIQMC step uses, no stored
 * matrix.  Requires a real symmetric H (every model except Transcorrelated1D) and one-word addresses.  This is the device
 * counterpart of multiplying a basis-ordered coefficient vector in the reference's exact diagonalisation
 * (ExactDiagonalization/basis_set_representation.jl:35-60; KrylovKit driver ext/KrylovKitExt.jl:23-46) with the matrix-free
 * mul! of DictVectors/pdvec.jl:810-822 as its operator. */
typedef struct rimu_sector rimu_sector;
int rimu_sector_create(rimu_ctx *ctx, const rimu_ham *ham, rimu_sector **out);
int rimu_sector_destroy(rimu_sector *s);
int rimu_sector_dim(const rimu_sector *s, uint64_t *dim_out);
/* rank of n addresses (host keys in, indices out; evaluated by the device code) and the addresses of a range of ranks */
int rimu_sector_rank(rimu_sector *s, const uint64_t *keys, int64_t n, int64_t *index_out);
int rimu_sector_keys(rimu_sector *s, int64_t first, int64_t count, uint64_t *keys_out);
/* dense vectors: device arrays of dim doubles owned by the library */
int rimu_sector_vec_create(rimu_sector *s, double **d_out);                 /* zero vector */
int rimu_sector_vec_destroy(rimu_sector *s, double *d);
int rimu_sector_vec_set(rimu_sector *s, double *d, const int64_t *index, const double *vals, int64_t n); /* zero, then d[index] = vals */
int rimu_sector_vec_get(rimu_sector *s, const double *d, int64_t first, int64_t count, double *out);
int rimu_sector_vec_gather(rimu_sector *s, const double *d, const int64_t *index, int64_t n, double *out);
/* y = H x (mul!, pdvec.jl:810-822, on the complete sector); *ms_out (optional): CUDA-event duration of the kernel */
int rimu_sector_mul(rimu_sector *s, const double *d_x, double *d_y, float *ms_out);
int rimu_sector_axpby(rimu_sector *s, double a, const double *d_x, double b, double *d_y);  /* y = a x + b y */
int rimu_sector_dot(rimu_sector *s, const double *d_x, const double *d_y, double *out);
/* conversions between the dictionary vector (Float64) and the dense layout */
int rimu_sector_from_vec(rimu_sector *s, rimu_vec *v, double *d_out);
int rimu_sector_to_vec(rimu_sector *s, const double *d, rimu_vec *v);

/* ---- the step ------------------------------------------------------------ */
int rimu_step(rimu_ctx *ctx, const rimu_ham *ham, const rimu_step_params *params,
              rimu_vec *src, rimu_vec *dst, rimu_step_stats *stats_out);

/* ---- a batch of steps ------------------------------------------------------
 * nsteps x { apply_operator!(wm, w, v, FirstOrderTransitionOperator(H, shift, dt)); swap(v, w); update_shift_parameters! }
 * = the body of advance!(::FCIQMC) (fciqmc.jl:126-181) with the shift strategies of
 * strategies_and_params/shiftstrategy.jl:77-213 and the abort rules of the reference (dead population, max_length, a strategy
 * that asks to stop) -- evaluated ON THE DEVICE, so that the kernels of step k+1 are enqueued before step k has finished.
 * A step of a small problem (the reference's own benchmark sizes, BASELINE config 1) is bound by launch latency and by the
 * host round trip of the walker number; a batch has neither.  Step k uses the Philox key of (params->seed, params->step + k):
 * the trajectory is the one nsteps rimu_step calls produce (the device evaluates the same shift formulas; its log() may
 * differ from the host's in the last bit).  One rank, partitioned method, not ordered: otherwise, and for vectors too
 * large to profit, the call simply runs step by step.  A chunk of steps that runs out of working memory is rolled back to
 * its snapshot and repeated step by step, so the call never returns a half-finished state. */
enum { RIMU_SHIFT_DONT_UPDATE = 0,                     /* shiftstrategy.jl:77-91 (stops once norm >= target_walkers) */
       RIMU_SHIFT_LOG_UPDATE = 1,                      /* :124-146 */
       RIMU_SHIFT_LOG_UPDATE_AFTER_TARGET = 2,         /* :100-122 */
       RIMU_SHIFT_DOUBLE_LOG_UPDATE = 3,               /* :160-181 */
       RIMU_SHIFT_DOUBLE_LOG_UPDATE_AFTER_TARGET = 4   /* :190-215 */ };
typedef struct {
    int32_t strategy;        /* RIMU_SHIFT_* */
    int32_t shift_mode;      /* in/out: DefaultShiftParameters.shift_mode (shiftstrategy.jl:32-38) */
    double target_walkers, zeta, xi;
    double shift, pnorm;     /* in/out: shift and previous walker number */
    int64_t max_length;      /* abort when the vector grows beyond this many entries (0 = no limit) */
} rimu_shift_params;
/* A frozen projector (FrozenDVec, DictVectors/projectors.jl:141-176): host (key, value) pairs.  rimu_advance evaluates
 * dot(projector, v) (pdvec.jl:773-779) after every step -- the reports of the ProjectedEnergy / Projector post-step
 * strategies (strategies_and_params/poststepstrategy.jl:50-121) -- without leaving the device. */
typedef struct {
    const uint64_t *keys;    /* n addresses, `words` uint64 each */
    const double *values;
    int64_t n;
} rimu_projector;
#define RIMU_MAX_PROJECTORS 8
/* v: current vector, w: scratch partner of the same type.  params->shift is ignored (sp->shift is the shift).  On return
 * *steps_done steps were taken (< nsteps only when the run ended: dead population, max_length, DontUpdate target reached --
 * the state is that of the last step taken, as in the reference), stats_out[k] / shift_out[k] (optional) hold the statistics
 * of step k and the shift AFTER its update, proj_out[k * nproj + j] = dot(projectors[j], v after step k), and *result_in_w
 * tells which vector holds the current state.  Both are written on every exit path: after an error (e.g. RIMU_ERR_WORKMEM in
 * step k) the first *steps_done steps stay taken.  w is scratch: whatever it held is overwritten. */
int rimu_advance(rimu_ctx *ctx, const rimu_ham *ham, const rimu_step_params *params, rimu_shift_params *sp,
                 rimu_vec *v, rimu_vec *w, int64_t nsteps, const rimu_projector *projectors, int32_t nproj,
                 rimu_step_stats *stats_out, double *shift_out, double *proj_out,
                 int64_t *steps_done, int32_t *result_in_w);
int rimu_sizeof_shift_params(void);

/* per-step Philox key derivation, exported so hosts/oracles can reproduce streams */
void rimu_step_key(uint64_t seed, uint64_t step, uint32_t key_out[2]);
void rimu_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* RIMU_B200_H */
