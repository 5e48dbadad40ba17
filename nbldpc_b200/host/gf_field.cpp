#include "gf_field.h"
#include "../../include/nbldpc_b200.h"
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

bool CGF::Initial(int GFq) {
    q = GFq; p = 0;
    while ((1 << p) < q) ++p;
    TableMultiply.assign((size_t)q * q, 0); TableInverse.assign(q, 0);
    if (nbl_gf_tables(q, 0, TableMultiply.data(), TableInverse.data()) != NBL_OK) {
        std::cerr << "GF(" << q << "): " << nbl_last_error() << std::endl;
        exit(-1);
    }
    std::ostringstream name;
    name << "./SRC/Arith.Table.GF." << q << ".txt";
    if (std::ifstream(name.str()).good() && nbl_gf_check_table_file(name.str().c_str(), q) != 0) {
        std::cerr << name.str() << " differs from the generated GF(" << q << ") tables: " << nbl_last_error() << std::endl;
        exit(-1);
    }
    return true;
}

int CGF::GFInverse(int a) const {
    if (a == 0) { std::cerr << "Div 0 Error!" << std::endl; exit(-1); }     // as GF.cpp:41-45
    return TableInverse[a];
}
