// Internal host-side structures shared by the loaders (code.cpp) and the CUDA decoder (decoder.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/nbldpc_b200.h"

struct nbl_code {
    int q = 0, p = 0, N = 0, M = 0, E = 0, max_dv = 0, max_dc = 0, prim_poly = 0;
    std::vector<int> mul, inv;                       // q*q, q
    // check side (row-major CSR, file order).  Edge id e = position in chk_var; this is the id used
    // for every message array on the device.
    std::vector<int> chk_ptr, chk_var, chk_coef, chk_dv;
    // variable side (column-major CSR, file order); var_edge maps a variable-side slot to its edge id
    std::vector<int> var_ptr, var_chk, var_coef, var_dc, var_edge;
};

void nbl_set_error(const std::string &msg);
int nbl_default_poly(int q);
