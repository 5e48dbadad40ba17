// Simulation driver: the frame loop of the reference's main.cpp:13-65 on the GPU decode path.
//   nbldpc_sim [profile] [--fp32 | --fp64] [--device D] [--per-frame] [--gpus G [--batch B]]
// profile: NBLDPC.Profile.txt by default, same format, same console table.
//   default      each round the `parallel` workers run their host stages (the reference's own source / noise generators) and are
//                decoded with ONE batched device call (CCommBatch) instead of the reference's parallel_for over CPU decoders;
//                counters are merged exactly as main.cpp:48-51 does.
//   --per-frame  the reference's per-frame Transmission() path (one device call per frame).
//   --gpus G     device Monte-Carlo on G GPUs of this box: every GPU simulates B frames per round (source, CRC, encoder, channel,
//                demapper, decoder and error counting all on the device, frames keyed by their global index), the 8 x u64 counter
//                vectors are summed with ONE NCCL all-reduce per round (nbl_simulate_sharded = main.cpp:46-51 + Comm.cpp:484-493)
//                and the reference's stop rule (Simulation.cpp:373-376) is applied to the totals.  Totals do not depend on G when
//                G * B is kept constant.
// Nothing here reads the environment.
#include "link_chain.h"
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

static void die(const char *what) {
    std::cerr << what << ": " << nbl_last_error() << std::endl;
    exit(-1);
}

// the --gpus path: one Eb/N0 point after the other on the device Monte-Carlo
static int run_device_monte_carlo(CSimulation &Simulation, int gpus, long long batch) {
    CComm Comm;                                   // worker 0: code, puncture set, constellation, sigma -- host side only
    Comm.Initial(Simulation, 0);
    CNBLDPC &C = Comm.NBLDPC;
    nbl_comm *comm = nullptr;
    if (nbl_comm_create(&comm, gpus, nullptr) != NBL_OK) die("nbl_comm_create");
    std::vector<nbl_decoder *> dec(gpus, nullptr);
    nbl_decoder_config cfg{};
    cfg.method = Simulation.decodeMethod; cfg.max_iter = Simulation.maxIter; cfg.precision = Simulation.precision; cfg.early_stop = 1;
    cfg.max_batch = (int)batch;
    if (C.PunctureLen > 0) cfg.tie_mode = NBL_TIES_REFERENCE;
    if (cfg.method == EMS_DECODE) { cfg.p0 = C.EMS_Nm; cfg.p1 = C.EMS_Nc; cfg.factor = C.EMS_Correction_Factor; cfg.offset = C.EMS_Correction_Offset; }
    else if (cfg.method == T_EMS_DECODE) { cfg.p0 = C.TEMS_Nr; cfg.p1 = C.TEMS_Nc; cfg.factor = C.TEMS_Factor; cfg.offset = C.TEMS_Offset; }
    else if (cfg.method == BS_TEMS_DECODE) { cfg.p0 = C.basic_set_Nm; cfg.p1 = C.basic_set_Nc; cfg.factor = C.basic_set_factor; cfg.offset = C.basic_set_offset; }
    else cfg.factor = 1;
    if (cfg.method == OSD_DECODE) { std::cerr << "--gpus: decode method 6 (OSD alone) runs on the batched path only" << std::endl; exit(-1); }
    for (int g = 0; g < gpus; ++g) {
        cfg.device = nbl_comm_device(comm, g);
        if (nbl_decoder_create(&dec[g], C.code(), &cfg, nullptr) != NBL_OK) die("decoder");
        if (C.OSD_Order >= 0) {
            nbl_osd_config oc{};
            oc.order = C.OSD_Order; oc.factor = C.OSD_Alpha; oc.flag = C.OSD_flag; oc.crc_len = C.CRCLen; oc.crc_correct_len = C.CRCrate;
            if (nbl_decoder_enable_osd(dec[g], &oc) != NBL_OK) die("OSD");
        }
    }
    std::vector<double> cons;
    if (Comm.modOrder != 2) for (int k = 0; k < Comm.modOrder; ++k) { cons.push_back(Comm.CONSTELLATION[k].Real); cons.push_back(Comm.CONSTELLATION[k].Image); }
    const double K = Comm.MSG_SYM_LEN, Kb = Comm.MSG_BIT_LEN;
    std::cout << "Device Monte-Carlo on " << gpus << " GPU(s), " << batch << " frames per GPU and round, kernel " << nbl_decoder_kernel(dec[0]) << std::endl;
    Simulation.Show(Screen_Head);
    while (Simulation.NextSNR()) {
        Simulation.ClearSimuCount();
        nbl_sim_config sc{};
        sc.ebn0_db = Simulation.EbN0; sc.modulation = Comm.modOrder; sc.constellation = cons.empty() ? nullptr : cons.data();
        sc.seed = (uint64_t)Simulation.randomseed; sc.first_frame = 0; sc.random_msg = Simulation.randomMsg; sc.crc_len = Simulation.crcLen;
        sc.punct_sym = C.PuncturePosition; sc.n_punct = C.PunctureLen;
        uint64_t tot[NBL_NUM_COUNTERS] = {0};
        // Simulation.cpp:373-376 with "frames so far" in the place of simCycle * parallel: go on while ANY criterion is unmet
        for (uint64_t round = 0; (double)tot[NBL_CNT_FRAMES] <= Simulation.minSimCycle || (double)tot[NBL_CNT_ERR_FRAMES] <= Simulation.minErrFrame ||
                                 (double)tot[NBL_CNT_UNDETECTED] <= Simulation.U_minErrFrame; ++round) {
            if (nbl_simulate_sharded(comm, dec.data(), &sc, (uint64_t)batch, round, tot) != NBL_OK) die("nbl_simulate_sharded");
            const double frames = (double)tot[NBL_CNT_FRAMES];
            Simulation.simCycle = frames / Simulation.parallel;
            Simulation.errFrame = (double)tot[NBL_CNT_ERR_FRAMES]; Simulation.errSym = (double)tot[NBL_CNT_ERR_SYMS]; Simulation.errBit = (double)tot[NBL_CNT_ERR_BITS];
            Simulation.U_errFrame = (double)tot[NBL_CNT_UNDETECTED];
            Simulation.sumIterations = (double)tot[NBL_CNT_ITERATIONS];
            Simulation.FER = Simulation.errFrame / frames; Simulation.SER = Simulation.errSym / (frames * K); Simulation.BER = Simulation.errBit / (frames * Kb);
            Simulation.U_FER = Simulation.U_errFrame / frames;
            Simulation.Show(Screen_Sim_Data);
        }
        Simulation.Show(Screen_Sim_End_Data);
        std::cout << "  frames " << tot[NBL_CNT_FRAMES] << "  mean iterations " << Simulation.sumIterations / (double)tot[NBL_CNT_FRAMES] << "  counters";
        for (int i = 0; i < NBL_NUM_COUNTERS; ++i) std::cout << ' ' << tot[i];
        std::cout << std::endl;
    }
    for (nbl_decoder *d : dec) nbl_decoder_destroy(d);
    nbl_comm_destroy(comm);
    std::cout << std::endl;
    return 0;
}

int main(int argc, char **argv) {
    const char *profile = "NBLDPC.Profile.txt";
    bool per_frame = false;
    int precision = -1, device = 0, gpus = 0;
    long long batch = 148 * 64;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](const char *flag) { if (i + 1 >= argc) { std::cerr << flag << " needs a value" << std::endl; exit(2); } return atoll(argv[++i]); };
        if (a == "--fp32") precision = 0;
        else if (a == "--fp64") precision = 1;
        else if (a == "--per-frame") per_frame = true;
        else if (a == "--device") device = (int)value("--device");
        else if (a == "--gpus") gpus = (int)value("--gpus");
        else if (a == "--batch") batch = value("--batch");
        else if (a.rfind("--", 0) == 0) { std::cerr << "usage: nbldpc_sim [profile] [--fp32|--fp64] [--device D] [--per-frame] [--gpus G [--batch B]]" << std::endl; return 2; }
        else profile = argv[i];
    }
    CSimulation Simulation;
    Simulation.Initial(profile);
    if (precision >= 0) Simulation.precision = precision;
    Simulation.device = device;
    Simulation.Show(Screen_Logo);
    Simulation.Show(Screen_Conf);
    if (gpus > 0) {
        if (batch < 1 || batch > (1 << 24)) { std::cerr << "--batch out of range" << std::endl; return 2; }
        return run_device_monte_carlo(Simulation, gpus, batch);
    }
    std::unique_ptr<CComm[]> Comm(new CComm[Simulation.parallel]);
    for (int i = 0; i < Simulation.parallel; ++i) Comm[i].Initial(Simulation, i);
    CCommBatch batchrun;
    Simulation.Show(Screen_Head);
    while (Simulation.NextSNR()) {
        Simulation.ClearSimuCount();
        for (int i = 0; i < Simulation.parallel; ++i) Comm[i].SetEbN0(Simulation, i);
        while (Simulation.SimulateThisSNR()) {
            if (per_frame) for (int i = 0; i < Simulation.parallel; ++i) Comm[i].Transmission();
            else batchrun.Run(Comm.get(), Simulation.parallel);
            for (int i = 0; i < Simulation.parallel; ++i) Comm[i].Err(Simulation);
            Simulation.Show(Screen_Sim_Data);
        }
        Simulation.Show(Screen_Sim_End_Data);
        std::cout << "  mean iterations " << Simulation.sumIterations / ((Simulation.simCycle - 1) * Simulation.parallel) << std::endl;
    }
    std::cout << std::endl;
    return 0;
}
