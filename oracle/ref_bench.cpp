// CPU-baseline timer (test/bench infrastructure, ours).  Links against the UNMODIFIED reference
// objects and runs the reference's own per-frame chain CComm::Transmission() (Comm.cpp:181-191)
// on T worker threads, exactly as main.cpp:27-51 does with `parallel` workers, but timed with a
// steady clock around the frame loop (the reference's clock() is CPU time on Linux).
//
// usage (cwd must hold the code/constellation files and ./SRC/ tables):
//     ref_bench <profile.txt> <threads> <frames_per_thread>
// prints one JSON line: {"frames":..,"seconds":..,"threads":..,"err_frames":..,"info_bits_per_frame":..}
#include "Simulation.h"
#include "Comm.h"
#include <chrono>
#include <thread>
#include <vector>
#include <cstdio>

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_bench profile threads frames_per_thread\n"); return 2; }
    int T = atoi(argv[2]), F = atoi(argv[3]);
    CSimulation sim;
    sim.Initial(argv[1]);
    sim.parallel = T;
    std::vector<CComm> comm(T);
    for (int i = 0; i < T; ++i) comm[i].Initial(sim, i);
    sim.NextSNR();
    sim.ClearSimuCount();
    for (int i = 0; i < T; ++i) comm[i].SetEbN0(sim, i);

    std::vector<double> errf(T, 0.0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int i = 0; i < T; ++i)
        pool.emplace_back([&, i] {
            for (int f = 0; f < F; ++f) {
                comm[i].Transmission();
                bool bad = false;
                for (int s = 0; s < comm[i].MSG_SYM_LEN; ++s) bad |= comm[i].TX_MSG_SYM[s] != comm[i].RX_MSG_SYM[s];
                errf[i] += bad;
            }
        });
    for (auto &t : pool) t.join();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double ef = 0; for (double e : errf) ef += e;
    printf("{\"frames\": %d, \"seconds\": %.6f, \"threads\": %d, \"err_frames\": %.0f, \"info_bits_per_frame\": %d, \"coded_bits_per_frame\": %d}\n",
           T * F, sec, T, ef, comm[0].MSG_BIT_LEN, comm[0].CODE_BIT_LEN);
    return 0;
}
