// stage_b.h -- host driver of the Stage-B device search (see stage_b.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <vector>
#include "../../include/ambigram_b200.h"

namespace ambi {

struct DevDag;

class StageB {
public:
    StageB();
    ~StageB();
    int init(cudaStream_t stream);
    void release();
    // first valid topological order for each DAG (orientation per reversed[i], flipped if none)
    int run(int count, const ambi_dag *dags, const int32_t *reversed, ambi_order_result *results);
    int run_all(const ambi_dag &dag, int forward, int32_t max_orders, std::vector<std::vector<int32_t>> &out,
                std::vector<std::vector<int32_t>> &bkp_out, unsigned long long *tested, bool *truncated,
                const int32_t *inv_src = nullptr, const int32_t *inv_tgt = nullptr, int32_t max_id = 0, int32_t path_stride = 0,
                std::vector<std::vector<int32_t>> *paths_out = nullptr);
    const char *error() const { return err_; }
    unsigned long long launches() const { return launches_; }
    double kernel_ms() const { return kernel_ms_; }        // device time of the last k_first_valid_order launch

private:
    int ensure(int count);
    int to_devdag(const ambi_dag &g, int reversed, DevDag &d, int index);
    cudaStream_t stream_ = nullptr;
    int cap_ = 0, sms_ = 148;
    void *d_dags_ = nullptr, *d_out_ = nullptr, *d_work_ = nullptr, *h_dags_ = nullptr, *h_out_ = nullptr;
    void *d_all_ = nullptr;            // --all scratch: counters | orders | lengths | paths (grow-only)
    size_t all_bytes_ = 0;
    unsigned long long launches_ = 0;
    cudaEvent_t ev_[2] = {nullptr, nullptr};
    double kernel_ms_ = 0;
    char err_[256];
};

}  // namespace ambi
