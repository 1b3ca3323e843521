// Stand-in for boost::thread_group / boost::bind (Boost 1.59) on top of <thread> / <functional>: the reference's row bands
// (src/PixelWisePyramid.cpp:424-436) run on real threads here too.  TEST INFRASTRUCTURE ONLY (see opencv2/opencv.hpp).
#pragma once
#include <functional>
#include <thread>
#include <vector>
namespace boost {
using std::bind;
class thread_group {
public:
    std::vector<std::thread> threads;
    template <class F> void create_thread(F f) { threads.emplace_back(f); }
    void join_all() { for (auto& t : threads) if (t.joinable()) t.join(); threads.clear(); }
    ~thread_group() { join_all(); }
};
}  // namespace boost
