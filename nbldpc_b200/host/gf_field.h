// GF(q) lookups with the reference's CGF surface (GF.h / GF.cpp:27-47).  Tables come from the C ABI
// (nbl_gf_tables, generated from the primitive polynomial) instead of ./SRC/Arith.Table.GF.q.txt; when that file
// is present it is cross-checked against the generated tables.
#pragma once
#include <vector>

class CGF {
public:
    int q = 0, p = 0;
    std::vector<int> TableMultiply, TableInverse;
    bool Initial(int GFq);
    int GFAdd(int a, int b) const { return a ^ b; }
    int GFMultiply(int a, int b) const { return TableMultiply[a * q + b]; }
    int GFInverse(int a) const;
};
