/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the NB-LDPC decode path.
 *
 * A plain-C, double-precision restatement of the reference decode path of YongonY/NBLDPC
 * (soft demapper, QSPA / EMS / T-EMS decoders).  It follows the reference's order of floating
 * point operations, so its outputs are bit-identical to the compiled reference (oracle/_ref) --
 * pinned by tests/golden/ (vectors dumped from the real reference) and by tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product path (nbldpc_b200/) never calls it and has no CPU fallback.
 */
#ifndef NBLDPC_ORACLE_H
#define NBLDPC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_code orc_code;

/* GF(q) tables from the primitive polynomial: add = XOR, mul = carry-less product mod poly
 * (what Arith.Table.GF.q.txt holds; reference lookups GF.cpp:27-47).  mul is q*q, inv is q. */
void orc_gf_tables(int q, int poly, int *mul, int *inv);

/* Code graph exactly as CNBLDPC::Initial holds it (NBLDPC.cpp:147-263): both edge lists in FILE
 * order, flattened with row stride max degree.  Cross indices VarLinkDc / ChkLinkDv are rebuilt. */
orc_code *orc_code_create(int q, int poly, int N, int M,
                          const int *var_deg, const int *chk_deg,
                          const int *var_link, const int *var_coef, int var_stride,
                          const int *chk_link, const int *chk_coef, int chk_stride);
void orc_code_destroy(orc_code *c);

/* CComm::Demodulate (Comm.cpp:340-407).  BPSK: y_re[N*p]; punct_bit may be NULL.
 * QAM (nQAM == q): y[2*N] interleaved re/im, constellation[2*q] interleaved.  Out: L_ch[N][q-1]. */
void orc_demap_bpsk(int q, int N, const double *y_re, double sigma, const int *punct_sym, int n_punct, double *L_ch);
void orc_demap_qam(int q, int N, const double *y, const double *constellation, double sigma,
                   const int *punct_sym, int n_punct, double *L_ch);

/* CNBLDPC::Decoding (NBLDPC.cpp:607-641) for method 1 (QSPA), 2 (EMS), 4 (T-EMS), 7 (BS-TEMS).
 * p0,p1 = (n_m,n_c) for EMS, (n_r,n_c) for T-EMS, (basic_set_Nm, basic_set_Nc) for BS-TEMS; factor/offset as in the profile.
 * L_ch, L_post: [N][q-1] doubles.  Returns 1 converged / 0 not (as the reference);
 * *iters = value of the reference's iter_number when it returned (maxIter on failure). */
int orc_decode(const orc_code *c, int method, int max_iter, int p0, int p1, double factor, double offset,
               const double *L_ch, int *out_sym, double *L_post, int *iters);

/* Bit-level ordered-statistics post-processing (OSD.h:5-199; matrices NBLDPC.cpp:380-457, OSD.h:440-510).
 * crc_len in {8, 16, 24} (anything else hangs the reference, OSD.h:472-499), crc_rate = crc_correctLen = number of CRC
 * parity rows stacked on H.  fix_matrepr = 0 reproduces the reference's binary image (the block of 1/alpha is zero,
 * GF.cpp:137), 1 uses the correct one. */
typedef struct orc_osd orc_osd;
orc_osd *orc_osd_create(const orc_code *c, int crc_len, int crc_rate, int fix_matrepr);
void orc_osd_destroy(orc_osd *o);
/* Decoding_OSD_bit.  sym [N]: in = decisions of the iterative decoder (ignored when flag), out = OSD decisions;
 * sym_sum = LLR_Sym_sum [N][q-1] (ignored when flag: the channel LLRs rank the bits). */
void orc_osd_run(orc_osd *o, int order, int flag, const double *L_ch, const double *sym_sum, int *sym);
void orc_osd_last(const orc_osd *o, int *order, int *min_distance);
/* orc_decode + the reference's "if (OSD_Order >= 0 && !decode_correct) Decoding_OSD_bit(...)" tail */
int orc_decode_osd(const orc_code *c, int method, int max_iter, int p0, int p1, double factor, double offset,
                   const double *L_ch, int *out_sym, double *L_post, int *iters,
                   orc_osd *osd, int osd_order, double osd_alpha, int osd_flag);

/* number of T-EMS configuration leaves visited in the last orc_decode call (tems_route_num, NBLDPC.cpp:1896) */
double orc_last_tems_leaves(void);
/* analysis switch, off by default: EMS configuration sums formed afresh instead of in the reference's add / subtract running sum */
void orc_set_ems_fresh_sums(int on);

#ifdef __cplusplus
}
#endif
#endif
