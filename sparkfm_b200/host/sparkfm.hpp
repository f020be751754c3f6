// sparkfm.hpp -- header-only C++ mirror of the reference's operator interface for the hot path,
// above the C ABI (include/sparkfm_b200.h).  The reference is JVM code and no JVM exists in the
// build image, so this is the compiled-language host layer; names, argument meaning and error
// behaviour follow the Scala (paths relative to src/main/scala/io/edstud/spark/):
//   Task                     Task.scala:3-6
//   DataSet                  DataSet.scala:9-62        (size, dimension = max index, cache)
//   FMModel                  fm/FMModel.scala:9-65     (predict, computeRMSE via Model.scala:13-19)
//   FMLearn                  fm/FMLearn.scala:10-16    (plugin boundary)
//   FM / learnWith           fm/FM.scala:15-33, fm/impl/FactorizationMachines.scala:30-51
// plus SGD (extends FMLearn) and FMWithSGD::train, which BASELINE.json north_star names.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sparkfm_b200.h"

namespace sparkfm {

enum class Task { Regression = SFM_TASK_REGRESSION, Classification = SFM_TASK_CLASSIFICATION };

struct Error : std::runtime_error {  // the Scala throws plain Exceptions (DataCollection.scala:36)
    int status;
    Error(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};

inline void check(int st, const sfm_handle* h = nullptr) {
    if (st == SFM_OK) return;
    std::string m = sfm_status_string(st);
    if (h) m += std::string(": ") + sfm_last_error(h);
    throw Error(st, m);
}

// RDD[(Double, SparseVector[Double])] packed as CSR; rows keep stored order and duplicates.
class DataSet {
public:
    std::string name;
    std::vector<float> labels;
    std::vector<int64_t> row_ptr{0};
    std::vector<int32_t> idx;
    std::vector<float> val;

    explicit DataSet(std::string n = "dataset") : name(std::move(n)) {}
    void add(double label, const std::vector<int32_t>& index, const std::vector<double>& data) {
        if (index.size() != data.size()) throw Error(SFM_ERR_ARG, "index/data length mismatch");
        labels.push_back((float)label);
        idx.insert(idx.end(), index.begin(), index.end());
        for (double x : data) val.push_back((float)x);
        row_ptr.push_back((int64_t)idx.size());
    }
    bool isEmpty() const { return labels.empty(); }                     // DataSet.scala:11
    int size() const { return isEmpty() ? 0 : (int)labels.size(); }      // :23-25
    int dimension() const {                                              // :27-29 (max index)
        if (isEmpty()) return 0;
        int m = INT32_MIN;
        for (size_t r = 0; r + 1 < row_ptr.size(); ++r) {
            if (row_ptr[r + 1] == row_ptr[r]) throw Error(SFM_ERR_ARG, "empty.max: row without features");
            for (int64_t j = row_ptr[r]; j < row_ptr[r + 1]; ++j) m = idx[j] > m ? idx[j] : m;
        }
        return m;
    }
};

class FMModel {
public:
    const int num_attribute, num_factor;
    double reg0 = 0, regw = 0, regv = 10;                                // fm/FMModel.scala:29-31

    FMModel(int num_attribute_, int num_factor_, Task task = Task::Regression, double init_mean = 0,
            double init_stdev = 0.01, uint64_t seed = 0, int device = 0, uint64_t sampler_seed = 42)
        : num_attribute(num_attribute_), num_factor(num_factor_) {
        sfm_config c{};
        c.abi_version = SFM_ABI_VERSION;
        c.task = (int)task;
        c.k = num_factor;
        c.k0 = c.k1 = 1;                                                  // :25-26
        c.device = device;
        c.n_slots = (int64_t)num_attribute + 1;                           // :18
        c.regv = 10.f;
        c.step_size = 0.1f;
        c.mini_batch_fraction = 1.f;
        c.sampler_seed = sampler_seed;
        check(sfm_create(&c, &h_));
        check(sfm_init_model(h_, init_mean, init_stdev, seed), h_);       // :19-22 (seeded here)
    }
    ~FMModel() { sfm_destroy(h_); }
    FMModel(const FMModel&) = delete;
    FMModel& operator=(const FMModel&) = delete;
    sfm_handle* handle() const { return h_; }

    // fm/FMModel.scala:34-55 for one sparse vector (a one-row batch on the device)
    double predict(const std::vector<int32_t>& index, const std::vector<double>& data) const {
        std::vector<float> x(data.begin(), data.end());
        const int64_t rp[2] = {0, (int64_t)index.size()};
        float out = 0.f;
        check(sfm_predict(h_, rp, index.data(), x.data(), 1, &out), h_);
        return out;
    }
    std::vector<float> predict(const DataSet& d) const {                  // rdd.mapValues(predict)
        std::vector<float> out(d.labels.size());
        check(sfm_predict(h_, d.row_ptr.data(), d.idx.data(), d.val.data(), (int64_t)d.labels.size(),
                          out.data()), h_);
        return out;
    }
    void cache(const DataSet& d) {                                        // DataSet.scala:50-54
        if (resident_ == &d) return;
        check(sfm_load_dataset(h_, d.row_ptr.data(), d.idx.data(), d.val.data(), d.labels.data(),
                               (int64_t)d.labels.size(), 0), h_);
        resident_ = &d;
    }
    double computeRMSE(const DataSet& d) {                                // Model.scala:13-19
        cache(d);
        double m[5];
        check(sfm_evaluate(h_, m), h_);
        return m[0];
    }
    void parameters(double* w0, std::vector<double>* w, std::vector<double>* v) const {
        w->resize((size_t)num_attribute + 1);
        v->resize(((size_t)num_attribute + 1) * num_factor);
        check(sfm_get_model_f64(h_, w0, w->data(), v->data()), h_);
    }
    void save(const std::string& path) const { check(sfm_save(h_, path.c_str()), h_); }

private:
    sfm_handle* h_ = nullptr;
    const DataSet* resident_ = nullptr;
};

class FMLearn {                                                           // fm/FMLearn.scala:10-16
public:
    virtual ~FMLearn() = default;
    virtual FMModel& learn(FMModel& fm, const DataSet& dataset) = 0;
};

class SGD : public FMLearn {
public:
    std::vector<double> lossHistory;
    SGD(double stepSize, double r0, double r1, double r2, double miniBatchFraction)
        : step_(stepSize), r0_(r0), r1_(r1), r2_(r2), frac_(miniBatchFraction) {}
    static std::unique_ptr<SGD> run(double stepSize = 0.1, double r0 = 0, double r1 = 0, double r2 = 0,
                                    double miniBatchFraction = 1.0) {     // cf. ALS.run, ALS.scala:202
        return std::unique_ptr<SGD>(new SGD(stepSize, r0, r1, r2, miniBatchFraction));
    }
    FMModel& learn(FMModel& fm, const DataSet& dataset) override {
        fm.cache(dataset);
        fm.reg0 = r0_; fm.regw = r1_; fm.regv = r2_;
        check(sfm_set_hyper(fm.handle(), (float)r0_, (float)r1_, (float)r2_, (float)step_, (float)frac_),
              fm.handle());
        double loss = 0;
        int64_t batch = 0;
        check(sfm_train_step(fm.handle(), nullptr, -1, ++iteration_, &loss, &batch), fm.handle());
        lossHistory.push_back(loss);
        return fm;
    }

private:
    double step_, r0_, r1_, r2_, frac_;
    int64_t iteration_ = 0;
};

// fm/lib/ALS.scala:11-208: the learner the reference ships.  One learn() = one sweep over w0, w, V
// with the model's own reg0 / regw / regv; refQuirks = bug-compatible with the reference
// (see sfm_als_sweep in include/sparkfm_b200.h).
class ALS : public FMLearn {
public:
    std::vector<double> rmseHistory;
    explicit ALS(bool refQuirks = false) : quirks_(refQuirks) {}
    static std::unique_ptr<ALS> run(bool refQuirks = false) {            // ALS.scala:202-208
        return std::unique_ptr<ALS>(new ALS(refQuirks));
    }
    FMModel& learn(FMModel& fm, const DataSet& dataset) override {
        fm.cache(dataset);
        sfm_config cfg;
        check(sfm_get_config(fm.handle(), &cfg), fm.handle());
        check(sfm_set_hyper(fm.handle(), (float)fm.reg0, (float)fm.regw, (float)fm.regv, cfg.step_size,
                            cfg.mini_batch_fraction), fm.handle());
        double rmse = 0;
        check(sfm_als_sweep(fm.handle(), quirks_ ? SFM_ALS_REF_QUIRKS : 0, &rmse), fm.handle());
        rmseHistory.push_back(rmse);
        return fm;
    }

private:
    bool quirks_;
};

// fm/FM.scala:25-33 + fm/impl/FactorizationMachines.scala:30-51
class FM {
public:
    FM(const DataSet& dataset, int numFactor, Task task = Task::Regression, int maxIteration = 100,
       int timeout = 0)
        : ds_(dataset), k_(numFactor), task_(task), iters_(maxIteration) { (void)timeout; }
    std::vector<double> rmseHistory;
    std::unique_ptr<FMModel> learnWith(FMLearn& fml, uint64_t seed = 0) {
        std::unique_ptr<FMModel> fm(new FMModel(ds_.dimension(), k_, task_, 0, 0.01, seed));
        for (int i = 1; i <= iters_; ++i) {
            rmseHistory.push_back(fm->computeRMSE(ds_));                  // :43 (logged upstream)
            fml.learn(*fm, ds_);                                          // :45 plugin boundary
        }
        return fm;
    }

private:
    const DataSet& ds_;
    int k_;
    Task task_;
    int iters_;
};

struct FMWithSGD {
    static std::unique_ptr<FMModel> train(const DataSet& input, Task task, int numIterations,
                                          double stepSize, double miniBatchFraction, int k,
                                          double r0, double r1, double r2, double initStd,
                                          std::vector<double>* lossHistory = nullptr) {
        std::unique_ptr<FMModel> fm(new FMModel(input.dimension(), k, task, 0, initStd, 0));
        SGD sgd(stepSize, r0, r1, r2, miniBatchFraction);
        for (int i = 0; i < numIterations; ++i) sgd.learn(*fm, input);
        if (lossHistory) *lossHistory = sgd.lossHistory;
        return fm;
    }
};

}  // namespace sparkfm
