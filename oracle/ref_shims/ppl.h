// Portability shim for the MSVC-only <ppl.h> that the reference driver includes
// (/root/reference/main.cpp:9,46).  Test infrastructure only.
#pragma once
#include <thread>
#include <vector>
namespace Concurrency {
template <class F> void parallel_for(int first, int last, F body) {
    std::vector<std::thread> pool;
    for (int i = first; i < last; ++i) pool.emplace_back([=] { body(i); });
    for (auto &t : pool) t.join();
}
}
