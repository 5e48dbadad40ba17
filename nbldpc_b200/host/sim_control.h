// Experiment control: same public surface as the reference's CSimulation (Simulation.h:31-120), written for
// this harness.  Reads the positional NBLDPC.Profile.txt format (Simulation.cpp:57-108), keeps the running
// counters that CComm::Err fills, iterates Eb/N0 and applies the reference's stop rule (Simulation.cpp:373-376).
#pragma once
#include <chrono>
#include <string>

#define BP_DECODE 1
#define EMS_DECODE 2
#define MinMax_DECODE 3
#define T_EMS_DECODE 4
#define T_MinMax_DECODE 5
#define OSD_DECODE 6
#define BS_TEMS_DECODE 7

#define Screen_Logo 0
#define Screen_Conf 1
#define Screen_Head 3
#define Screen_Sim_Data 7
#define Screen_Sim_End_Data 8

class CSimulation {
public:
    std::string ProfileFileName, NonBinaryFileName, ConstellationFileName;
    int GFq = 0, decodeMethod = 0, parallel = 1, maxIter = 0, randomMsg = 0;
    int PuntureVarDegree = 0, nQAM = 2, randomseed = 1;
    double snrBegin = 0, snrStep = 1, snrStop = 0, EbN0 = 0;
    int crcLen = 0, crc_correctLen = 0;
    int OSD_order = -1; double OSD_factor = 1; int OSD_flag = 0;
    int ems_nm = 0, ems_nc = 0; double ems_factor = 1, ems_offset = 0;
    int tems_nr = 0, tems_nc = 0; double tems_factor = 1, tems_offset = 0;
    int bs_tems_nm = 0, bs_tems_nc = 0; double bs_tems_factor = 1, bs_tems_offset = 0;
    double errFrame = 0, errBit = 0, errSym = 0, FER = 0, BER = 0, SER = 0;
    double U_errFrame = 0, U_errBit = 0, U_errSym = 0, U_FER = 0, U_BER = 0, U_SER = 0;
    double simCycle = 0;
    int minSimCycle = 0, minErrFrame = 0, U_minErrFrame = 0, showSimFrameStep = 1;
    // additions of this harness (not in the reference): decoded-iteration statistics and GPU precision
    double sumIterations = 0;
    int precision = 1;      // NBL_PREC_F64 by default: bit-exact decisions with the reference (driver flag --fp32 / --fp64)
    int device = 0;         // CUDA device of the per-frame / batched paths (driver flag --device)
    std::chrono::steady_clock::time_point start;

    int Initial(std::string profilename);
    int Show(int mode);
    int ClearSimuCount(void);
    bool NextSNR(void);
    bool SimulateThisSNR(void);
};
