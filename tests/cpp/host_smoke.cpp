// C++ host-layer check: drives the C ABI through sparkfm.hpp exactly like the reference's
// driver drives FM(...).learnWith(ALS.run) (driver.scala:106-112) and prints the numbers the
// pytest wrapper compares with the Python / oracle path.
#include <cstdio>
#include <cstdlib>

#include "../../sparkfm_b200/host/sparkfm.hpp"

static uint64_t mix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

int main() {
    using namespace sparkfm;
    try {
        DataSet ds("cpp.train");
        const int n_rows = 3000, n_feat = 500;
        for (int r = 0; r < n_rows; ++r) {
            std::vector<int32_t> idx;
            std::vector<double> val;
            const int m = 3 + (int)(mix64(r) % 9);
            for (int j = 0; j < m; ++j) {
                idx.push_back((int32_t)(mix64(r * 131 + j) % n_feat));   // duplicates may occur
                val.push_back(1.0 + (double)(mix64(r * 977 + j) % 4) * 0.25);
            }
            ds.add((mix64(r ^ 0xABCD) & 1) ? 1.0 : -1.0, idx, val);
        }
        FM fm(ds, 8, Task::Classification, 5);
        auto sgd = SGD::run(0.3, 0.0, 1e-4, 1e-3, 0.5);
        auto model = fm.learnWith(*sgd, 7);
        std::printf("dimension %d size %d\n", ds.dimension(), ds.size());
        for (size_t i = 0; i < sgd->lossHistory.size(); ++i)
            std::printf("iter %zu rmse_before %.9g loss %.9g\n", i + 1, fm.rmseHistory[i], sgd->lossHistory[i]);
        std::printf("predict0 %.9g\n", model->predict({ds.idx[0], ds.idx[1]}, {1.0, 2.0}));
        // error behaviour: an out-of-range index is a reported error, not UB
        try {
            model->predict({n_feat + 5}, {1.0});
            std::printf("ERROR: no exception\n");
            return 1;
        } catch (const Error& e) {
            std::printf("index error status %d\n", e.status);
        }
        return 0;
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
}
