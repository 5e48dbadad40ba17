/* nbldpc_b200 -- C ABI of the B200-native non-binary LDPC decode path.
 *
 * Drop-in boundary for the decode path of YongonY/NBLDPC.  The reference has no FFI; its boundary is the
 * public C++ surface  CNBLDPC::Initial / CNBLDPC::Decoding / CComm::Demodulate / CComm::Err.  Every entry
 * point below names the reference interface it replaces (file:line into the reference tree).  Plain
 * pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success and a negative nbl_status on error; nbl_last_error() gives text.
 *     Nothing throws across the ABI and nothing calls exit() (the reference exit(-1)s on config errors,
 *     e.g. NBLDPC.cpp:282-286, 619-638; the C++ harness on top keeps that behaviour).
 *   - "LLR vector" layout is the reference's: q-1 doubles per symbol, entry k-1 = ln P(x=k)/P(x=0)
 *     (NBLDPC.cpp:1542-1562, Comm.cpp:360-398).  Batched arrays are dense [B][N][q-1].
 *   - data pointers may be HOST or DEVICE pointers (detected with cudaPointerGetAttributes); device
 *     pointers must belong to the decoder's device.  Host data is staged through device buffers on the
 *     decoder's stream (pass pinned host memory for asynchronous copies).
 *   - handles are not thread-safe; distinct decoder handles are independent (one CUDA stream each).
 *   - there is no CPU fallback: every compute entry point fails with NBL_ERR_CUDA when no B200 is present.
 */
#ifndef NBLDPC_B200_H
#define NBLDPC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NBL_ABI_VERSION 4

typedef enum {
    NBL_OK = 0,
    NBL_ERR_ARG = -1,     /* bad argument / unsupported configuration */
    NBL_ERR_IO = -2,      /* file missing or malformed */
    NBL_ERR_CUDA = -3,    /* CUDA runtime error (message in nbl_last_error) */
    NBL_ERR_ALLOC = -4
} nbl_status;

/* decode methods: same numbers as the reference's profile / Simulation.h:3-9 */
#define NBL_METHOD_QSPA 1 /* BP_DECODE    -> CNBLDPC::Decoding_BP   NBLDPC.cpp:643-776  (FWHT check node)   */
#define NBL_METHOD_EMS 2  /* EMS_DECODE   -> CNBLDPC::Decoding_EMS  NBLDPC.cpp:778-927                      */
#define NBL_METHOD_TEMS 4 /* T_EMS_DECODE -> CNBLDPC::Decoding_TEMS NBLDPC.cpp:929-1143                     */
#define NBL_METHOD_BSTEMS 7 /* BS_TEMS_DECODE -> CNBLDPC::Decoding_BS_TEMS NBLDPC.cpp:1145-1396 (basic-set T-EMS) */

#define NBL_PREC_F32 0 /* fast mode: fp32 messages and arithmetic (QSPA transform in fp64)               */
#define NBL_PREC_F64 1 /* parity mode: fp64 everywhere, reference operation order, no FMA contraction   */

#define NBL_TIES_AUTO 0
#define NBL_TIES_FAST 1
#define NBL_TIES_REFERENCE 2
#define NBL_TIES_BITEXACT 3

/* nbl_decoder_config.kernel: which iteration kernel family serves the decoder.  AUTO picks the fastest one the code and
 * the parameters allow (what nbl_decoder_kernel reports); the other two exist for A/B measurements and for the tests that
 * compare the kernels with each other -- results are identical by construction.                                          */
#define NBL_KERNEL_AUTO 0
#define NBL_KERNEL_GROUP 1        /* the generic q-threads-per-row kernel (any code, any parameters)                    */
#define NBL_KERNEL_WARP_GENERIC 2 /* the warp-per-row kernels in their degree-generic instantiation (degree guards kept) */

typedef struct nbl_code nbl_code;       /* GF(q) tables + parity-check graph, resident on one device */
typedef struct nbl_decoder nbl_decoder; /* message workspace for up to max_batch frames + a stream    */
typedef struct nbl_encoder nbl_encoder; /* systematic encoder of one code on one device                 */

int nbl_abi_version(void);
const char *nbl_last_error(void);
int nbl_device_count(void); /* number of visible CUDA devices (0 without a GPU; never an error) */

/* ---- code and table loading: replaces CGF::Initial (GF.cpp:50-155) and the graph part of
 *      CNBLDPC::Initial (NBLDPC.cpp:140-263).  These are host-only until a decoder is created. ---- */

/* GF(q) arithmetic from the primitive polynomial of Arith.Table.GF.q.txt (0 = the polynomial the
 * reference ships for that q: 7,11,19,37,67,137,285,529).  mul is q*q ints, inv q ints (inv[0] = 0,
 * where the reference's GFInverse(0) exits, GF.cpp:41-45). */
int nbl_gf_tables(int q, int prim_poly, int *mul, int *inv);
/* cross-check an Arith.Table.GF.q.txt file against the generated tables (0 = identical). */
int nbl_gf_check_table_file(const char *path, int q);

/* graph from explicit edge lists, both in the reference's file order (chk_* row-major CSR,
 * var_* column-major CSR; var_ptr may be NULL => variable side ordered by ascending check index). */
int nbl_code_create(nbl_code **out, int q, int prim_poly, int N, int M,
                    const int *chk_ptr, const int *chk_var, const int *chk_coef,
                    const int *var_ptr, const int *var_chk, const int *var_coef);
/* the reference's alist-like code file (divsalar.*, BDS.*; parsed at NBLDPC.cpp:147-205) */
int nbl_code_load_file(nbl_code **out, const char *path, int prim_poly);
/* binary image H_bit_GF*.txt (no reference loader exists; layout in DESIGN.md) */
int nbl_code_load_hbit(nbl_code **out, const char *path, int q, int prim_poly);
void nbl_code_destroy(nbl_code *code);

typedef struct {
    int q, p, N, M, E, max_dv, max_dc, prim_poly;
} nbl_code_info;
int nbl_code_get_info(const nbl_code *code, nbl_code_info *info);
/* copy the edge lists out (any pointer may be NULL): chk_ptr[M+1], chk_var[E], chk_coef[E], chk_dv[E]
 * (= ChkLinkDv), var_ptr[N+1], var_chk[E], var_coef[E], var_dc[E] (= VarLinkDc) */
int nbl_code_get_graph(const nbl_code *code, int *chk_ptr, int *chk_var, int *chk_coef, int *chk_dv,
                       int *var_ptr, int *var_chk, int *var_coef, int *var_dc);
/* variables whose degree equals punct_degree (the reference's puncture rule, NBLDPC.cpp:163-177);
 * returns the count, fills positions[] (capacity N) when not NULL. */
int nbl_code_puncture_set(const nbl_code *code, int punct_degree, int *positions);

/* ---- decoder: replaces the per-method scratch of CNBLDPC::Initial (NBLDPC.cpp:267-329) and
 *      CNBLDPC::Decoding + Decoding_BP/EMS/TEMS (NBLDPC.cpp:607-1143) ---- */
typedef struct {
    int method;      /* NBL_METHOD_*                                                           */
    int max_iter;    /* maxIter of the profile                                                 */
    int p0;          /* EMS: n_m      T-EMS: n_r      BS-TEMS: basic_set_Nm      QSPA: ignored */
    int p1;          /* EMS: n_c      T-EMS: n_c      BS-TEMS: basic_set_Nc      QSPA: ignored */
    double factor;   /* EMS / T-EMS / BS-TEMS correction factor (divide)                       */
    double offset;   /* EMS / T-EMS / BS-TEMS offset (soft threshold)                          */
    int precision;   /* NBL_PREC_*                                                             */
    int early_stop;  /* 1: frames leave the batch when their syndrome is zero (reference rule);
                        0: every frame runs max_iter iterations, outputs are still the reference's
                           (decision / iteration count latched at the first zero syndrome)       */
    int max_batch;   /* frames per nbl_decode_batch call that the workspace is sized for        */
    int device;      /* CUDA device ordinal                                                     */
    int tie_mode;    /* T-EMS, exactly tied configuration costs (punctured symbols, quantised LLRs):
                        NBL_TIES_REFERENCE: the recorded path is the one the reference's depth-first
                        enumeration keeps (NBLDPC.cpp:1892-1944) -- on the one-warp-per-row kernel too (n_r <= 2: a second
                        pass ranks the configurations of exactly equal cost; about 1/2 of the FAST throughput), else
                        on the group kernel;
                        NBL_TIES_FAST: first in the fast kernel's scan order (identical results whenever no
                        two configurations of a syndrome tie exactly, i.e. for continuous LLRs);
                        NBL_TIES_AUTO (0): REFERENCE in the fp64 parity mode, FAST in the fp32 mode;
                        NBL_TIES_BITEXACT: a verification mode.  The reference forms a configuration's cost in ONE
                        running sum (+= going down, -= coming back), whose rounding residue decides between
                        configurations that are equal in exact arithmetic -- which T-EMS produces by itself
                        after two iterations, because outputs are copies and sums of inputs.  This mode replays
                        that search sequentially (one thread per check row) and so reproduces the reference's
                        messages bit for bit in fp64 for any n_c; orders of magnitude slower                 */
    int kernel;      /* NBL_KERNEL_* (0 = AUTO).  Nothing in the library reads the environment.              */
} nbl_decoder_config;

/* stream: a cudaStream_t (as void*) to launch on, or NULL for a private non-blocking stream */
int nbl_decoder_create(nbl_decoder **out, const nbl_code *code, const nbl_decoder_config *cfg, void *stream);
void nbl_decoder_destroy(nbl_decoder *dec);
/* which iteration kernel this decoder launches (static string): "tems_warp", "bstems_warp", "ems_warp", "qspa_warp"
 * (one warp per check row, q >= 32, d_c <= 4), "tems_gf16", "bstems_gf16", "ems_gf16", "qspa_gf16" (four lanes per row), or "group" (the generic
 * q-threads-per-row kernel: any code, any parameters, several times slower).  For logs and tests: a configuration that
 * silently moved to "group" is a performance regression, not an error. */
const char *nbl_decoder_kernel(const nbl_decoder *dec);

/* Bit-level ordered-statistics post-processing of the frames the iterative decoder fails on -- replaces
 * CNBLDPC::Decoding_OSD_bit (OSD.h:5-199) and the "if (OSD_Order >= 0 && !decode_correct)" tail of every Decoding_*
 * (NBLDPC.cpp:770-775, 921-925, 1135-1139).  Fields are the reference profile's (Simulation.cpp:57-108):
 * once enabled, nbl_decode_batch / nbl_demap_decode_* / nbl_simulate replace the symbols of every frame whose syndrome
 * never became zero by the OSD decision; `converged` stays 0 for them, as the reference's return value does.           */
typedef struct {
    int order;            /* OSD_order 0..3 (number of simultaneously flipped reliable bits); < 0 switches OSD off      */
    double factor;        /* OSD_factor: LLR_Sym_sum = factor * LLR_Sym_sum + L_post at every iteration                 */
    int flag;             /* OSD_flag 1: start from the channel's hard decisions and rank by the channel LLRs           */
    int crc_len;          /* crcLen: 8, 16 or 24 (the reference does not start with another value)           */
    int crc_correct_len;  /* crc_correctLen: CRC parity rows stacked on H (0..crc_len)                                  */
    int fix_matrepr;      /* 0: the reference's binary image of H, in which the block of the field element 1/alpha is
                             all-zero because Mat.Repr is read one record short (GF.cpp:137); 1: correct one */
} nbl_osd_config;
int nbl_decoder_enable_osd(nbl_decoder *dec, const nbl_osd_config *cfg);
/* decode method 6 (NBLDPC.cpp:625-627): OSD alone, from the channel LLRs, on every frame.  L_ch [B][N][q-1], out [B][N] */
int nbl_osd_batch(nbl_decoder *dec, const double *L_ch, int B, int *out_sym);
void *nbl_decoder_stream(nbl_decoder *dec);
int nbl_decoder_sync(nbl_decoder *dec);

/* CNBLDPC::Decoding for B frames (B = 1 is the reference's call, Comm.cpp:417).
 *   L_ch       [B][N][q-1] doubles (channel LLRs, as RX_LLR_SYM)
 *   out_sym    [B][N] ints        hard decisions (as DecodeOutput / RX_DECODE_SYM)
 *   converged  [B] ints or NULL   the reference's return value, 1 = syndrome satisfied
 *   iterations [B] ints or NULL   value of the reference's iter_number at return (max_iter on failure)
 *   L_post     [B][N][q-1] doubles or NULL   posterior at return (public member CNBLDPC::L_post)
 * The call is synchronous for host outputs; with device outputs it only enqueues on the stream. */
int nbl_decode_batch(nbl_decoder *dec, const double *L_ch, int B, int *out_sym, int *converged,
                     int *iterations, double *L_post);

/* ---- channel front end: replaces CComm::Demodulate (Comm.cpp:340-407); sigma from CComm::SetEbN0
 *      (Comm.cpp:157-179).  Output is L_ch [B][N][q-1] doubles (host or device). ---- */
/* BPSK: y_re [B][(N - n_punct) * p] received real parts in transmit order */
int nbl_demap_bpsk(nbl_decoder *dec, const double *y_re, int B, double sigma, const int *punct_sym,
                   int n_punct, double *L_ch);
/* nQAM == q: y [B][N - n_punct][2] received points, constellation [q][2] */
int nbl_demap_qam(nbl_decoder *dec, const double *y, int B, const double *constellation, double sigma,
                  const int *punct_sym, int n_punct, double *L_ch);
/* fused front end + decode (what CComm::Demodulate followed by CComm::Decode does per frame,
 * Comm.cpp:188-189): received samples in, decisions out, channel LLRs never leave the device.
 * modulation: 2 = BPSK (y is [B][(N-n_punct)*p]), q = QAM (y is [B][N-n_punct][2]) */
int nbl_demap_decode_batch(nbl_decoder *dec, int modulation, const double *y, int B, const double *constellation,
                           double sigma, const int *punct_sym, int n_punct, int *out_sym, int *converged,
                           int *iterations);

/* ---- systematic encoder: replaces CNBLDPC::InitialEncode (NBLDPC.cpp:477-560, 1474-1538: Gauss elimination with
 *      the reference's pivot / row-swap / column-swap rules, on the host, once) and CNBLDPC::Encode
 *      (NBLDPC.cpp:562-604) for B frames on the device.  Fails with "NB matrix is not full rank" where the
 *      reference prints that and exits. ---- */
int nbl_encoder_create(nbl_encoder **out, const nbl_code *code, int device, void *stream);
void nbl_encoder_destroy(nbl_encoder *enc);
/* K = N - M message symbols per frame; n_swaps = column swaps the elimination recorded (0 for every shipped code) */
int nbl_encoder_get_info(const nbl_encoder *enc, int *K, int *n_swaps);
/* msg [B][K] ints (as TX_MSG_SYM), code_sym [B][N] ints (as TX_CODE_SYM); host or device pointers.  Like the
 * reference, the message a receiver recovers is code_sym[f][0..K-1] (NBLDPC.cpp:599-602 writes it back into
 * msg_sym; identical to msg unless columns were swapped). */
int nbl_encode_batch(nbl_encoder *enc, const int *msg, int B, int *code_sym);

/* ---- device Monte-Carlo: the frame loop of main.cpp:43-55 (source -> map -> AWGN -> demap -> decode
 *      -> CComm::Err, Comm.cpp:446-503) for `frames` frames starting at global index first_frame.
 *      Noise is counter-based (Philox4x32-10 keyed by (seed, frame index)) so that counters do not
 *      depend on how frames are split over calls, streams or GPUs. ---- */
#define NBL_CNT_FRAMES 0
#define NBL_CNT_ERR_FRAMES 1
#define NBL_CNT_ERR_SYMS 2
#define NBL_CNT_ERR_BITS 3
#define NBL_CNT_CONVERGED 4   /* decoder returned 1                                         */
#define NBL_CNT_UNDETECTED 5  /* crc_len == 0: converged to a wrong codeword; crc_len > 0: the reference's U_errFrame --
                                 message in error AND its CRC check passes (Comm.cpp:462-469)  */
#define NBL_CNT_ITERATIONS 6  /* sum of iteration counts                                    */
#define NBL_CNT_CRC_OK 7      /* crc_len > 0: frames whose decoded message passes the CRC check */
#define NBL_NUM_COUNTERS 8

typedef struct {
    double ebn0_db;
    int modulation;             /* 2 = BPSK, q = QAM                                                 */
    const double *constellation;/* [q][2] host pointer, QAM only                                     */
    uint64_t seed;
    uint64_t first_frame;       /* global index of the first frame of this call                      */
    uint64_t frames;            /* number of frames to simulate (any size; processed in max_batch chunks) */
    const int *codewords;       /* host [n_codewords][N] pool of valid codewords (NULL = all-zero codeword);
                                   frame f transmits codeword f % n_codewords                        */
    int n_codewords;
    int random_msg;             /* with n_codewords == 0: 1 = every frame draws its own K message symbols on the
                                   device (Philox keyed by (seed, frame index)) and encodes them there
                                   (the reference's randomMsg = 1, Comm.cpp:214-249 + CNBLDPC::Encode); 0 = all-zero */
    int crc_len;                /* crcLen 0, 8, 16 or 24: with random_msg the last crc_len message bits are the CRC of the
                                   others (CComm::CRCEncode, Comm.cpp:506-560), and the decoded message of every frame
                                   is CRC-checked for NBL_CNT_UNDETECTED / NBL_CNT_CRC_OK (CComm::CrcCheck, :562-630;
                                   like the reference: CRC24A on encode but CRC24B on check)                  */
    const int *punct_sym;       /* host [n_punct] punctured symbol positions (nbl_code_puncture_set; the reference's
                                   PunctureVarDegree rule, NBLDPC.cpp:163-177): not transmitted, all-zero LLRs at the decoder,
                                   sigma from R = K / (N - n_punct) as Comm.cpp:98.  NULL / 0: nothing punctured          */
    int n_punct;
} nbl_sim_config;

/* counters: host uint64[NBL_NUM_COUNTERS], ACCUMULATED into (caller zeroes). */
int nbl_simulate(nbl_decoder *dec, const nbl_sim_config *cfg, uint64_t *counters);
/* same, but leaves the sums in a device-resident uint64[8] (for an in-place NCCL all-reduce) and
 * does not synchronise; dev_counters is accumulated into.  All work is queued on the decoder's stream
 * (nbl_decoder_stream): zero, reduce or read the counters on THAT stream, or after nbl_decoder_sync -- a decoder created
 * with stream == NULL owns a non-blocking stream that the legacy default stream does not wait for. */
int nbl_simulate_async(nbl_decoder *dec, const nbl_sim_config *cfg, uint64_t *dev_counters);

/* ---- multi-GPU: frames shard by global frame index (nbl_sim_config.first_frame); the only collective of the path is the
 *      sum of the 8 x u64 counter vectors.  Replaces the reference's serial merge of its workers' counters
 *      (main.cpp:46-51 parallel_for + Comm.cpp:484-493 CComm::Err adding into CSimulation), so that every rank applies the
 *      stop rule (Simulation.cpp:373-376) to the same totals.  NCCL is bound at run time; world == 1 needs none. ---- */
typedef struct nbl_comm nbl_comm;
/* one process driving n_devices GPUs (devices == NULL: 0..n-1): ncclCommInitAll */
int nbl_comm_create(nbl_comm **out, int n_devices, const int *devices);
/* one process per GPU: rank 0 calls nbl_comm_unique_id and hands the 128 bytes to the other ranks by its own means
 * (file, MPI, torch.distributed broadcast); every rank then calls nbl_comm_create_rank (ncclCommInitRank) */
int nbl_comm_unique_id(void *id128);
int nbl_comm_create_rank(nbl_comm **out, int world, int rank, const void *id128, int device);
void nbl_comm_destroy(nbl_comm *comm);
int nbl_comm_size(const nbl_comm *comm);        /* ranks in the communicator */
int nbl_comm_local_ranks(const nbl_comm *comm); /* ranks this process holds: n_devices, or 1 */
int nbl_comm_rank(const nbl_comm *comm, int local_index);   /* global rank of local rank i (-1: out of range) */
int nbl_comm_device(const nbl_comm *comm, int local_index); /* its CUDA device */
/* In-place sum over all ranks of dev_counters (device uint64[NBL_NUM_COUNTERS], as filled by nbl_simulate_async), queued on
 * each decoder's stream behind the simulation it just ran: decoders[i] / dev_counters[i] belong to local rank i
 * (nbl_comm_local_ranks entries; the decoder must live on that rank's device).  Asynchronous: read the result after
 * nbl_decoder_sync or on the same stream.  One ncclAllReduce(sum, uint64 x 8) per rank, grouped. */
int nbl_allreduce_counters(nbl_comm *comm, nbl_decoder *const *decoders, uint64_t *const *dev_counters);

/* One round of a sharded Monte-Carlo run = one pass of the reference's frame loop main.cpp:43-55 (its `parallel` workers, then
 * the merge of their counters): local rank i (global rank r_i) simulates frames_per_rank frames starting at
 * cfg->first_frame + (round_index * world + r_i) * frames_per_rank on decoders[i] (cfg->frames is ignored), the per-rank counter
 * vectors are summed with nbl_allreduce_counters, and the round's total is ADDED to totals (host uint64[8]) on every rank.
 * Synchronous.  The union of the frames of rounds 0..R-1 does not depend on the number of ranks when
 * world * frames_per_rank is kept constant, hence neither do the totals. */
int nbl_simulate_sharded(nbl_comm *comm, nbl_decoder *const *decoders, const nbl_sim_config *cfg, uint64_t frames_per_rank,
                         uint64_t round_index, uint64_t *totals);

/* timing hook: milliseconds the GPU spent in the decode kernels of the last nbl_decode_batch /
 * nbl_demap_decode_batch / nbl_simulate call on this decoder (CUDA events on the decoder's stream),
 * and the number of kernel launches it made. */
int nbl_decoder_last_timing(nbl_decoder *dec, float *kernel_ms, int *launches);
/* per-launch duration (ms) of the fused iteration kernel in the last decode on this decoder: ms[i] is
 * iteration i+1.  Synchronises the stream.  Returns the number of entries written (<= capacity). */
int nbl_decoder_iter_timing(nbl_decoder *dec, float *ms, int capacity);
/* frames that had not yet returned (the reference returns a frame at its first zero syndrome, NBLDPC.cpp:1016-1026) when each
 * iteration of the last decode on this decoder ran: active[i] is iteration i+1.  With early_stop = 1 the iteration kernels work on
 * these frames only; beside nbl_decoder_iter_timing this shows what the tail iterations cost per active frame.  Synchronises. */
int nbl_decoder_iter_active(nbl_decoder *dec, int *active, int capacity);

#ifdef __cplusplus
}
#endif
#endif
