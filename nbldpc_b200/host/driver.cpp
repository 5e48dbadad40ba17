// Simulation driver: the frame loop of the reference's main.cpp:13-65 on the GPU decode path.
//   nbldpc_sim [profile]            NBLDPC.Profile.txt by default, same format, same console table
// Each round the `parallel` workers run their host stages and are decoded with ONE batched device call
// (CCommBatch) instead of the reference's parallel_for over CPU decoders; counters are merged exactly as
// main.cpp:48-51 does.  NBLDPC_PER_FRAME=1 switches to the reference's per-frame Transmission() path.
#include "link_chain.h"
#include <cstdlib>
#include <iostream>
#include <memory>

int main(int argc, char **argv) {
    CSimulation Simulation;
    Simulation.Initial(argc > 1 ? argv[1] : "NBLDPC.Profile.txt");
    Simulation.Show(Screen_Logo);
    Simulation.Show(Screen_Conf);
    std::unique_ptr<CComm[]> Comm(new CComm[Simulation.parallel]);
    for (int i = 0; i < Simulation.parallel; ++i) Comm[i].Initial(Simulation, i);
    const bool per_frame = getenv("NBLDPC_PER_FRAME") != nullptr;
    CCommBatch batch;
    Simulation.Show(Screen_Head);
    while (Simulation.NextSNR()) {
        Simulation.ClearSimuCount();
        for (int i = 0; i < Simulation.parallel; ++i) Comm[i].SetEbN0(Simulation, i);
        while (Simulation.SimulateThisSNR()) {
            if (per_frame) for (int i = 0; i < Simulation.parallel; ++i) Comm[i].Transmission();
            else batch.Run(Comm.get(), Simulation.parallel);
            for (int i = 0; i < Simulation.parallel; ++i) Comm[i].Err(Simulation);
            Simulation.Show(Screen_Sim_Data);
        }
        Simulation.Show(Screen_Sim_End_Data);
        std::cout << "  mean iterations " << Simulation.sumIterations / ((Simulation.simCycle - 1) * Simulation.parallel) << std::endl;
    }
    std::cout << std::endl;
    return 0;
}
