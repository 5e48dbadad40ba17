// Host-side loaders: GF(q) tables, parity-check graph from the reference's text formats.
// Replaces CGF::Initial (GF.cpp:50-155) and the parsing / cross-index part of CNBLDPC::Initial
// (NBLDPC.cpp:140-263) of the reference; written from the file formats, not from that code.
#include "internal.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

static thread_local std::string g_err;
void nbl_set_error(const std::string &msg) { g_err = msg; }
extern "C" const char *nbl_last_error(void) { return g_err.c_str(); }
extern "C" int nbl_abi_version(void) { return NBL_ABI_VERSION; }

int nbl_default_poly(int q) {
    switch (q) {  // primitive polynomials in the headers of the shipped Arith.Table.GF.q.txt files
        case 4: return 7; case 8: return 11; case 16: return 19; case 32: return 37;
        case 64: return 67; case 128: return 137; case 256: return 285; case 512: return 529;
    }
    return 0;
}

static bool valid_q(int q) { return q >= 4 && q <= 256 && (q & (q - 1)) == 0; }

extern "C" int nbl_gf_tables(int q, int prim_poly, int *mul, int *inv) {
    if (!(q >= 2 && q <= 512 && (q & (q - 1)) == 0) || !mul || !inv) { nbl_set_error("nbl_gf_tables: bad q"); return NBL_ERR_ARG; }
    if (!prim_poly) prim_poly = nbl_default_poly(q);
    // degree exactly log2 q: bit q set and nothing above it -- otherwise the reduction below would leave v >= q and index
    // past the tables (a polynomial can come from an untrusted Arith.Table header)
    if (!(prim_poly & q) || (prim_poly & ~(2 * q - 1)) || prim_poly < 0) { nbl_set_error("nbl_gf_tables: the polynomial must have degree log2(q)"); return NBL_ERR_ARG; }
    // powers of x: antilog/log tables, then products by index addition
    std::vector<int> alog(2 * q), lg(q, -1);
    int v = 1;
    for (int i = 0; i < q - 1; ++i) {
        if (lg[v] != -1) { nbl_set_error("nbl_gf_tables: polynomial is not primitive"); return NBL_ERR_ARG; }
        alog[i] = v; lg[v] = i;
        v <<= 1;
        if (v & q) v ^= prim_poly;
    }
    if (v != 1) { nbl_set_error("nbl_gf_tables: polynomial is not primitive"); return NBL_ERR_ARG; }
    for (int i = q - 1; i < 2 * q; ++i) alog[i] = alog[i - (q - 1)];
    for (int a = 0; a < q; ++a)
        for (int b = 0; b < q; ++b) mul[a * q + b] = (a && b) ? alog[lg[a] + lg[b]] : 0;
    inv[0] = 0;
    for (int a = 1; a < q; ++a) inv[a] = alog[(q - 1 - lg[a]) % (q - 1)];
    return NBL_OK;
}

extern "C" int nbl_gf_check_table_file(const char *path, int q) {
    std::ifstream f(path);
    if (!f) { nbl_set_error(std::string("cannot open ") + path); return NBL_ERR_IO; }
    std::string header, w;
    std::getline(f, header);                       // "GF(q) with Primitive Polynomial: P."
    int poly = 0;
    size_t colon = header.rfind(':');
    if (colon != std::string::npos) poly = atoi(header.c_str() + colon + 1);
    std::vector<int> mul(q * q), inv(q);
    int rc = nbl_gf_tables(q, poly, mul.data(), inv.data());
    if (rc) return rc;
    auto bad = [&](const char *what) { nbl_set_error(std::string(path) + ": " + what + " table differs"); return 1; };
    int x;
    f >> w >> w;                                   // "Multiply Table:"
    for (int i = 0; i < q * q; ++i) { if (!(f >> x) || x != mul[i]) return bad("multiply"); }
    f >> w >> w;                                   // "Add Table:"
    for (int i = 0; i < q * q; ++i) { if (!(f >> x) || x != ((i / q) ^ (i % q))) return bad("add"); }
    f >> w >> w;                                   // "Inverse Table:"
    for (int i = 0; i < q; ++i) { if (!(f >> x) || x != inv[i]) return bad("inverse"); }
    return 0;
}

// ---------------------------------------------------------------------------------------------
static int finish_code(nbl_code *c) {
    // cross indices (VarLinkDc / ChkLinkDv of the reference) and the edge-id map of the variable side
    c->E = (int)c->chk_var.size();
    c->chk_dv.assign(c->E, -1); c->var_dc.assign(c->E, -1); c->var_edge.assign(c->E, -1);
    c->max_dv = c->max_dc = 0;
    for (int r = 0; r < c->M; ++r) c->max_dc = std::max(c->max_dc, c->chk_ptr[r + 1] - c->chk_ptr[r]);
    for (int v = 0; v < c->N; ++v) c->max_dv = std::max(c->max_dv, c->var_ptr[v + 1] - c->var_ptr[v]);
    for (int v = 0; v < c->N; ++v)
        for (int s = c->var_ptr[v]; s < c->var_ptr[v + 1]; ++s) {
            int r = c->var_chk[s];
            if (r < 0 || r >= c->M) { nbl_set_error("code: check index out of range"); return NBL_ERR_IO; }
            for (int e = c->chk_ptr[r]; e < c->chk_ptr[r + 1]; ++e)
                if (c->chk_var[e] == v) {
                    if (c->chk_coef[e] != c->var_coef[s]) { nbl_set_error("code: edge coefficient differs between the two lists"); return NBL_ERR_IO; }
                    c->var_dc[s] = e - c->chk_ptr[r]; c->var_edge[s] = e; c->chk_dv[e] = s - c->var_ptr[v];
                }
            if (c->var_edge[s] < 0) { nbl_set_error("code: variable-side edge missing on the check side"); return NBL_ERR_IO; }
        }
    for (int e = 0; e < c->E; ++e) {
        if (c->chk_dv[e] < 0) { nbl_set_error("code: check-side edge missing on the variable side"); return NBL_ERR_IO; }
        if (c->chk_coef[e] <= 0 || c->chk_coef[e] >= c->q) { nbl_set_error("code: coefficient out of range"); return NBL_ERR_IO; }
    }
    if (c->max_dc > 16 || c->max_dc < 2) { nbl_set_error("code: check degree must be in 2..16"); return NBL_ERR_ARG; }
    return NBL_OK;
}

static int init_field(nbl_code *c, int q, int prim_poly) {
    if (!valid_q(q)) { nbl_set_error("unsupported field size (need q = 2^p, 4 <= q <= 256)"); return NBL_ERR_ARG; }
    c->q = q; c->p = 0; while ((1 << c->p) < q) ++c->p;
    c->prim_poly = prim_poly ? prim_poly : nbl_default_poly(q);
    c->mul.resize(q * q); c->inv.resize(q);
    return nbl_gf_tables(q, c->prim_poly, c->mul.data(), c->inv.data());
}

extern "C" int nbl_code_create(nbl_code **out, int q, int prim_poly, int N, int M, const int *chk_ptr, const int *chk_var,
                               const int *chk_coef, const int *var_ptr, const int *var_chk, const int *var_coef) {
    if (!out || !chk_ptr || !chk_var || !chk_coef || N <= 0 || M <= 0) { nbl_set_error("nbl_code_create: bad argument"); return NBL_ERR_ARG; }
    if (var_ptr && (!var_chk || !var_coef)) { nbl_set_error("nbl_code_create: var_ptr without var_chk / var_coef"); return NBL_ERR_ARG; }
    if (chk_ptr[0] != 0) { nbl_set_error("nbl_code_create: chk_ptr must start at 0"); return NBL_ERR_ARG; }
    for (int r = 0; r < M; ++r) if (chk_ptr[r + 1] < chk_ptr[r]) { nbl_set_error("nbl_code_create: chk_ptr must be non-decreasing"); return NBL_ERR_ARG; }
    {   // a (row, variable) pair may appear once: two edges between the same nodes would alias one message slot
        std::vector<int> seen(N, -1);
        for (int r = 0; r < M; ++r)
            for (int e = chk_ptr[r]; e < chk_ptr[r + 1]; ++e) {
                const int v = chk_var[e];
                if (v < 0 || v >= N) { nbl_set_error("nbl_code_create: variable index out of range"); return NBL_ERR_ARG; }
                if (seen[v] == r) { nbl_set_error("nbl_code_create: duplicate edge (same check row and variable twice)"); return NBL_ERR_ARG; }
                seen[v] = r;
            }
    }
    nbl_code *c = new nbl_code;
    int rc = init_field(c, q, prim_poly);
    if (rc) { delete c; return rc; }
    c->N = N; c->M = M;
    int E = chk_ptr[M];
    c->chk_ptr.assign(chk_ptr, chk_ptr + M + 1);
    c->chk_var.assign(chk_var, chk_var + E);
    c->chk_coef.assign(chk_coef, chk_coef + E);
    for (int e = 0; e < E; ++e) if (chk_var[e] < 0 || chk_var[e] >= N) { nbl_set_error("nbl_code_create: variable index out of range"); delete c; return NBL_ERR_ARG; }
    if (var_ptr) {
        c->var_ptr.assign(var_ptr, var_ptr + N + 1);
        bool mono = var_ptr[0] == 0;
        for (int v = 0; v < N; ++v) mono = mono && var_ptr[v + 1] >= var_ptr[v];
        if (!mono) { nbl_set_error("nbl_code_create: var_ptr must start at 0 and be non-decreasing"); delete c; return NBL_ERR_ARG; }
        if (var_ptr[N] != E) { nbl_set_error("nbl_code_create: edge counts differ"); delete c; return NBL_ERR_ARG; }
        c->var_chk.assign(var_chk, var_chk + E);
        c->var_coef.assign(var_coef, var_coef + E);
    } else {                                   // ascending check order (true of every shipped code file)
        c->var_ptr.assign(N + 1, 0);
        for (int e = 0; e < E; ++e) c->var_ptr[chk_var[e] + 1]++;
        for (int v = 0; v < N; ++v) c->var_ptr[v + 1] += c->var_ptr[v];
        c->var_chk.resize(E); c->var_coef.resize(E);
        std::vector<int> fill(c->var_ptr.begin(), c->var_ptr.end() - 1);
        for (int r = 0; r < M; ++r)
            for (int e = chk_ptr[r]; e < chk_ptr[r + 1]; ++e) {
                int s = fill[chk_var[e]]++;
                c->var_chk[s] = r; c->var_coef[s] = chk_coef[e];
            }
    }
    rc = finish_code(c);
    if (rc) { delete c; return rc; }
    *out = c;
    return NBL_OK;
}

extern "C" int nbl_code_load_file(nbl_code **out, const char *path, int prim_poly) {
    std::ifstream f(path);
    if (!f) { nbl_set_error(std::string("cannot open code file ") + path); return NBL_ERR_IO; }
    int N, M, q, mdv, mdc;
    if (!(f >> N >> M >> q >> mdv >> mdc) || N <= 0 || M <= 0) { nbl_set_error("code file: bad header"); return NBL_ERR_IO; }
    std::vector<int> vdeg(N), cdeg(M);
    for (int &d : vdeg) if (!(f >> d) || d < 0) { nbl_set_error("code file: bad variable degree"); return NBL_ERR_IO; }
    for (int &d : cdeg) if (!(f >> d) || d < 0) { nbl_set_error("code file: bad check degree"); return NBL_ERR_IO; }
    std::vector<int> vptr(N + 1, 0), cptr(M + 1, 0);
    for (int v = 0; v < N; ++v) vptr[v + 1] = vptr[v] + vdeg[v];
    for (int r = 0; r < M; ++r) cptr[r + 1] = cptr[r] + cdeg[r];
    if (vptr[N] != cptr[M]) { nbl_set_error("code file: degree sums differ"); return NBL_ERR_IO; }
    int E = vptr[N];
    std::vector<int> vchk(E), vcoef(E), cvar(E), ccoef(E);
    for (int s = 0; s < E; ++s) { if (!(f >> vchk[s] >> vcoef[s])) { nbl_set_error("code file: truncated"); return NBL_ERR_IO; } vchk[s]--; }
    for (int e = 0; e < E; ++e) { if (!(f >> cvar[e] >> ccoef[e])) { nbl_set_error("code file: truncated"); return NBL_ERR_IO; } cvar[e]--; }
    int extra;
    if (f >> extra) { nbl_set_error("code file: trailing tokens"); return NBL_ERR_IO; }
    return nbl_code_create(out, q, prim_poly, N, M, cptr.data(), cvar.data(), ccoef.data(), vptr.data(), vchk.data(), vcoef.data());
}

extern "C" int nbl_code_load_hbit(nbl_code **out, const char *path, int q, int prim_poly) {
    if (!valid_q(q)) { nbl_set_error("nbl_code_load_hbit: bad q"); return NBL_ERR_ARG; }
    std::ifstream f(path);
    if (!f) { nbl_set_error(std::string("cannot open ") + path); return NBL_ERR_IO; }
    std::vector<std::vector<int>> B;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ls(line);
        std::vector<int> row; int x;
        while (ls >> x) row.push_back(x);
        if (!row.empty()) B.push_back(row);
    }
    if (B.empty()) { nbl_set_error("H_bit file is empty"); return NBL_ERR_IO; }
    for (auto &r : B) if (r.size() != B[0].size()) { nbl_set_error("H_bit file: ragged rows"); return NBL_ERR_IO; }
    int p = 0; while ((1 << p) < q) ++p;
    if (!prim_poly) prim_poly = nbl_default_poly(q);
    // degree exactly log2 q: bit q set and nothing above it -- otherwise the reduction below would leave v >= q and index
    // past the tables (a polynomial can come from an untrusted Arith.Table header)
    if (!(prim_poly & q) || (prim_poly & ~(2 * q - 1)) || prim_poly < 0) { nbl_set_error("nbl_gf_tables: the polynomial must have degree log2(q)"); return NBL_ERR_ARG; }
    // Orientation from the shape: more rows than columns = transposed image B[p*sym+l][p*chk+k];
    // otherwise H_bin[p*chk+k][p*sym+l].  Either way block(sym,chk)[l][k] = bit k of g*x^l.
    bool transposed = B.size() > B[0].size();
    int rows = (int)B.size(), cols = (int)B[0].size();
    int N = (transposed ? rows : cols) / p, M = (transposed ? cols : rows) / p;
    if (rows % p || cols % p) { nbl_set_error("H_bit file: shape is not a multiple of log2 q"); return NBL_ERR_IO; }
    auto bit = [&](int sym, int l, int chk, int k) { return transposed ? B[p * sym + l][p * chk + k] : B[p * chk + k][p * sym + l]; };
    std::vector<int> cptr(M + 1, 0), cvar, ccoef;
    for (int r = 0; r < M; ++r) {
        for (int s = 0; s < N; ++s) {
            int g = 0;
            for (int k = 0; k < p; ++k) g |= bit(s, 0, r, k) << k;
            if (!g) {
                for (int l = 0; l < p; ++l) for (int k = 0; k < p; ++k) if (bit(s, l, r, k)) { nbl_set_error("H_bit file: block is not a field-element matrix"); return NBL_ERR_IO; }
                continue;
            }
            int v = g;
            for (int l = 0; l < p; ++l) {
                for (int k = 0; k < p; ++k) if (bit(s, l, r, k) != ((v >> k) & 1)) { nbl_set_error("H_bit file: block is not a field-element matrix"); return NBL_ERR_IO; }
                v <<= 1; if (v & q) v ^= prim_poly;
            }
            cvar.push_back(s); ccoef.push_back(g);
        }
        cptr[r + 1] = (int)cvar.size();
    }
    return nbl_code_create(out, q, prim_poly, N, M, cptr.data(), cvar.data(), ccoef.data(), nullptr, nullptr, nullptr);
}

extern "C" void nbl_code_destroy(nbl_code *code) { delete code; }

extern "C" int nbl_code_get_info(const nbl_code *c, nbl_code_info *info) {
    if (!c || !info) { nbl_set_error("nbl_code_get_info: null"); return NBL_ERR_ARG; }
    *info = nbl_code_info{c->q, c->p, c->N, c->M, c->E, c->max_dv, c->max_dc, c->prim_poly};
    return NBL_OK;
}

extern "C" int nbl_code_get_graph(const nbl_code *c, int *chk_ptr, int *chk_var, int *chk_coef, int *chk_dv,
                                  int *var_ptr, int *var_chk, int *var_coef, int *var_dc) {
    if (!c) { nbl_set_error("nbl_code_get_graph: null"); return NBL_ERR_ARG; }
    auto cp = [](int *dst, const std::vector<int> &src) { if (dst) memcpy(dst, src.data(), src.size() * sizeof(int)); };
    cp(chk_ptr, c->chk_ptr); cp(chk_var, c->chk_var); cp(chk_coef, c->chk_coef); cp(chk_dv, c->chk_dv);
    cp(var_ptr, c->var_ptr); cp(var_chk, c->var_chk); cp(var_coef, c->var_coef); cp(var_dc, c->var_dc);
    return NBL_OK;
}

extern "C" int nbl_code_puncture_set(const nbl_code *c, int punct_degree, int *positions) {
    if (!c) { nbl_set_error("nbl_code_puncture_set: null"); return NBL_ERR_ARG; }
    int n = 0;
    for (int v = 0; v < c->N; ++v)
        if (c->var_ptr[v + 1] - c->var_ptr[v] == punct_degree) { if (positions) positions[n] = v; ++n; }
    return n;
}
