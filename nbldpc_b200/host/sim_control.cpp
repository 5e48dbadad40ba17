#include "sim_control.h"
#include <cstdlib>
#include <fstream>
#include <iostream>

// one value preceded by `labels` throw-away tokens (only the token COUNT matters, Simulation.cpp:57-108)
template <class V> static void field(std::ifstream &f, int labels, V &value) {
    std::string skip;
    for (int i = 0; i < labels; ++i) f >> skip;
    f >> value;
}

int CSimulation::Initial(std::string profilename) {
    ProfileFileName = profilename;
    std::ifstream f(profilename);
    if (!f.is_open()) { std::cerr << "Cannot open " << profilename << std::endl; exit(-1); }
    field(f, 1, GFq); field(f, 2, NonBinaryFileName); field(f, 2, PuntureVarDegree);
    field(f, 2, decodeMethod); field(f, 2, maxIter); field(f, 1, parallel);
    field(f, 1, crcLen); field(f, 1, crc_correctLen);
    field(f, 1, OSD_order); field(f, 1, OSD_factor); field(f, 1, OSD_flag);
    field(f, 2, ems_nm); field(f, 2, ems_nc); field(f, 2, ems_factor); field(f, 2, ems_offset);
    field(f, 2, tems_nr); field(f, 2, tems_nc); field(f, 2, tems_factor); field(f, 2, tems_offset);
    field(f, 2, bs_tems_nm); field(f, 2, bs_tems_nc); field(f, 2, bs_tems_factor); field(f, 2, bs_tems_offset);
    field(f, 2, snrBegin); field(f, 2, snrStep); field(f, 2, snrStop);
    field(f, 1, nQAM); field(f, 1, ConstellationFileName);
    field(f, 2, randomMsg); field(f, 3, minErrFrame); field(f, 3, U_minErrFrame); field(f, 3, minSimCycle);
    field(f, 2, randomseed); field(f, 3, showSimFrameStep);
    if (!f) { std::cerr << "Profile " << profilename << " is incomplete" << std::endl; exit(-1); }
    EbN0 = snrBegin - snrStep;
    return 0;
}

int CSimulation::Show(int mode) {
    using std::cout; using std::endl;
    switch (mode) {
    case Screen_Logo:
        cout << "******************************************************************************\n"
             << "*******   NBLDPC decoding simulation -- B200 decode path (nbldpc_b200)   *******\n"
             << "******************************************************************************\n" << endl;
        break;
    case Screen_Conf:
        cout << "Configuration lists as follows:\nCode: " << NonBinaryFileName << "\nGF: " << GFq
             << "\tPuncture Variable Degree: " << PuntureVarDegree << "\tMaximum iterations: " << maxIter
             << "\tparallel: " << parallel << endl;
        cout << "crcLen: " << crcLen << "\tcrcusedforcorrect: " << crc_correctLen << "\tOSD_order: " << OSD_order
             << "\tOSD_factor: " << OSD_factor << "\tOSD_flag: " << OSD_flag << endl;
        if (decodeMethod == BP_DECODE) cout << "Algorithm: BP" << decodeMethod << endl;
        else if (decodeMethod == EMS_DECODE)
            cout << "Algorithm: EMS\tNm: " << ems_nm << "\tNc: " << ems_nc << "\tFactor: " << ems_factor << "\tOffset: " << ems_offset << endl;
        else if (decodeMethod == T_EMS_DECODE)
            cout << "Algorithm: Trellis EMS\tNr: " << tems_nr << "\tNc: " << tems_nc << "\tFactor: " << tems_factor << "\tOffset: " << tems_offset << endl;
        else cout << "Algorithm: " << decodeMethod << endl;
        cout << "Modulation: " << nQAM << "-QAM\tConf: " << ConstellationFileName << endl;
        cout << "SNR: " << snrBegin << ":" << snrStep << ":" << snrStop << (randomMsg ? "\tRandom Sequence\n" : "\tALL 0 sequence\n")
             << "Min Err Frame: " << minErrFrame << "\tMin U_Err Frame: " << U_minErrFrame << "\tMin Sim Frame: " << minSimCycle
             << "\tRand Seed: " << randomseed << "\tGPU precision: " << (precision ? "fp64 (parity)" : "fp32 (fast)") << endl;
        break;
    case Screen_Head:
        cout << "EbN0\tError\tCRCmiss\tBER\t\tSER\t\tFER\t\tU-FER\t\tTime" << endl;
        break;
    case Screen_Sim_Data:
        if (long(simCycle) % showSimFrameStep == 0) {
            cout << std::defaultfloat << EbN0 << '\t' << errFrame << '\t' << U_errFrame << '\t';
            cout << std::scientific << BER << '\t' << SER << '\t' << FER << '\t' << U_FER << '\t' << simCycle << '\r';
        }
        break;
    case Screen_Sim_End_Data: {
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();   // wall time
        cout << std::defaultfloat << EbN0 << '\t' << errFrame << '\t' << U_errFrame << '\t';
        cout << std::scientific << BER << '\t' << SER << '\t' << FER << '\t' << U_FER << '\t' << std::defaultfloat << sec << endl;
        break;
    }
    default: break;
    }
    return 0;
}

int CSimulation::ClearSimuCount(void) {
    simCycle = 0;
    errFrame = errBit = errSym = 0;
    U_errFrame = U_errBit = U_errSym = 0;
    sumIterations = 0;
    start = std::chrono::steady_clock::now();
    return 0;
}

bool CSimulation::NextSNR(void) {
    EbN0 += snrStep;
    return !(EbN0 > snrStop);
}

// continue while ANY criterion is unmet; -1 disables a criterion
bool CSimulation::SimulateThisSNR(void) {
    return ((simCycle++) * parallel <= minSimCycle) || (errFrame <= minErrFrame) || (U_errFrame <= U_minErrFrame);
}
