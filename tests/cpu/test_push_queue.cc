// Host-thread model of the device-side work queue of finish_pair_and_push (atlas_b200/csrc/fourier.cu): the blocks of a sharded
// direct Fourier launch as tasks on a thread pool, the same sequence of atomic operations (arrival counter per latitude pair, slot
// reservation + publication by the completing block, chunk claims, head advance).  Checks the property the exchange relies on:
// every 128-row chunk of every completed pair is shipped exactly once, whatever the interleaving.  (The CUDA code itself runs in
// the emulated-rank GPU tests; this pins the protocol.)
#include <atomic>
#include <cstdio>
#include <random>
#include <thread>
#include <vector>
int main() {
    const int npairs = 97, nblk = 23, rows_per = 128;
    std::mt19937 rng(7);
    int bad = 0;
    for (int trial = 0; trial < 300; ++trial) {
        std::vector<int> L(npairs);
        for (auto& l : L) l = rng() % 1280;
        std::vector<std::atomic<int>> pair_done(npairs), pairs(npairs), claim(npairs);
        std::atomic<int> tail{0}, head{0};
        for (int i = 0; i < npairs; ++i) pair_done[i] = 0, pairs[i] = 0, claim[i] = 0;
        std::vector<std::vector<std::atomic<int>>> pushed(npairs);
        for (int p = 0; p < npairs; ++p) {
            const int nch = (2 * (L[p] + 1) + rows_per - 1) / rows_per;
            pushed[p] = std::vector<std::atomic<int>>(nch);
            for (auto& x : pushed[p]) x = 0;
        }
        auto block = [&](int pair) {
            const int done = pair_done[pair].fetch_add(1);
            if (done == nblk - 1) {
                const int slot = tail.fetch_add(1);
                pairs[slot].exchange(pair + 1);
            }
            for (;;) {
                int found = -1, chunk = 0;
                int s = head.load();
                const int t = tail.load();
                while (s < t) {
                    int pp = pairs[s].load();
                    while (pp == 0) pp = pairs[s].load();
                    const int nch = (2 * (L[pp - 1] + 1) + rows_per - 1) / rows_per;
                    const int c = claim[s].fetch_add(1);
                    if (c < nch) { found = pp - 1; chunk = c; break; }
                    int h = head.load();
                    while (h < s + 1 && !head.compare_exchange_weak(h, s + 1)) {}
                    ++s;
                }
                if (found < 0) break;
                pushed[found][chunk].fetch_add(1);
            }
        };
        std::vector<int> order;
        for (int p = 0; p < npairs; ++p)
            for (int b = 0; b < nblk; ++b) order.push_back(p);
        // blocks of a pair start close together, like the block list of a launch, with some shuffling
        for (size_t i = 0; i + 40 < order.size(); i += 7) std::swap(order[i], order[i + rng() % 40]);
        const int nthreads = 8 + trial % 25;
        std::atomic<size_t> next{0};
        std::vector<std::thread> th;
        for (int w = 0; w < nthreads; ++w)
            th.emplace_back([&] {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= order.size()) break;
                    block(order[i]);
                }
            });
        for (auto& t : th) t.join();
        for (int p = 0; p < npairs; ++p)
            for (auto& x : pushed[p])
                if (x.load() != 1) ++bad;
    }
    printf(bad ? "FAILED: %d chunks not shipped exactly once\n" : "ALL OK (%d)\n", bad);
    return bad != 0;
}
