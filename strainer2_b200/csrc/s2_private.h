// s2_private.h - structures shared by the device-side translation units (s2_capi.cu, s2_ingest.cu)
#pragma once
#include "../../include/strainer2_b200.h"
#include "s2_internal.h"
#include "s2_kernels.cuh"

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            s2_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,   \
                         cudaGetErrorString(e_));                                                  \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

#define CKN(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            s2_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,   \
                         cudaGetErrorString(e_));                                                  \
            return nullptr;                                                                        \
        }                                                                                          \
    } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
enum LaneState { LANE_FREE = 0, LANE_HELD = 1, LANE_INFLIGHT = 2 };

struct Lane {
    cudaStream_t stream = nullptr;
    uint8_t *h_buf = nullptr;          // pinned
    uint8_t *d_buf = nullptr;
    cudaEvent_t k0 = nullptr, k1 = nullptr;   // bracket the scan kernel on this lane's stream
    // scratch of the two-phase (partitioned) scan, allocated on first use and grown on demand
    uint64_t *part_pool = nullptr; uint64_t part_entries = 0;
    unsigned long long *part_cursor = nullptr; uint32_t *part_overflow = nullptr;
    LaneState state = LANE_FREE;
    bool timed = false;
    uint64_t seq = 0;                  // submission order, to recycle the oldest first
};

struct s2_ctx {
    int device = 0, n_sm = 0;
    uint64_t serial = 0;               // unique per s2_init: a thread's ingest pipeline recognises a context that is gone
    uint64_t batch_bytes = 0;
    int n_lanes = 0;
    std::vector<Lane> lanes;
    std::mutex mu;
    uint64_t next_seq = 1;
    unsigned long long *d_stats = nullptr;     // [0] hits [1] valid windows, accumulated on device
    unsigned long long *h_stats = nullptr;     // pinned mirror
    int grid_count = 0, grid_detect = 0;
    double kernel_ms = 0.0;
    uint64_t kernel_launches = 0;
    // enqueue-only scans on lane 0 (device-resident batches): one event pair per launch, harvested at sync
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
    cudaEvent_t user_ev[4] = { nullptr, nullptr, nullptr, nullptr };
    // grow-only scratch of s2_scan_detect (no cudaMalloc per call)
    struct Scratch { void *p = nullptr; size_t cap = 0; } det[6];
    // GPU ingest pipelines of this context (s2_ingest.cu): a small pool shared by all calling threads
    void *ingest_pool = nullptr;
};

static inline int scratch_reserve(s2_ctx::Scratch &s, size_t bytes)
{
    if (bytes <= s.cap) return 0;
    if (s.p) cudaFree(s.p);
    s.p = nullptr; s.cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    CK(cudaMalloc(&s.p, want));
    s.cap = want;
    return 0;
}


struct s2_table {
    s2_ctx *ctx = nullptr;
    S2TableView v = {};
    uint32_t *rank_slot = nullptr;     // first-occurrence rank -> slot
    uint32_t *rank_pos = nullptr;      // first-occurrence rank -> byte offset of that first window in the build stream
    uint32_t *scratch = nullptr;       // n_keys uint32 staging for fetch / store
    uint64_t n_keys = 0;
    bool partitioned = false;          // fingerprints do not fit L2: count scans go through the two-phase kernels
};


// s2_shutdown: the context's ingest pipelines go with it (s2_ingest.cu)
void s2_ingest_ctx_closing(s2_ctx *c);

// the count scan of one batch on one lane's stream (direct or two-phase), caller holds c->mu
int s2_launch_count_on_lane(s2_ctx *c, Lane &l, const uint8_t *d_bases, uint64_t n_bytes, s2_table *t, int col);
