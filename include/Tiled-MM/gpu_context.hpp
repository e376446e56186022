// Drop-in counterpart of reference src/Tiled-MM/gpu_context.hpp.  The reference's gpu_context OWNS n streams, n cuBLAS
// handles and one result stream (gpu_context.cpp:6-18); here those belong to the tmm_context behind mm_handle (1 H2D,
// 1 D2H, high-priority phase-1 chains, low-priority column-block streams, a pooled set of events - csrc/tmm_context.cu),
// and gpu_context is a view of them with the reference's accessors.
#pragma once
#include "../tiled_mm_b200.h"
#include "device_event.hpp"
#include "device_stream.hpp"
#include "gpu_blas_api.hpp"
#include "gpu_blas_handle.hpp"
#include "gpu_runtime_api.hpp"
#include "util.hpp"

#include <memory>
#include <stdexcept>
#include <vector>

namespace gpu {

class gpu_context {
public:
    explicit gpu_context(tmm_context* ctx) : ctx_(ctx) {}

    // stream_id in [0, n): the context's compute streams (gpu_context.cpp:24-31)
    runtime_api::StreamType get_stream(int stream_id) const { return static_cast<runtime_api::StreamType>(checked(TMM_STREAM_COMPUTE, stream_id)); }
    blas_api::HandleType get_blas_handle(int stream_id) const { return checked(TMM_STREAM_COMPUTE, stream_id); }
    device_stream& get_device_stream(int stream_id) {
        const runtime_api::StreamType s = get_stream(stream_id);  // throws when out of range
        if ((int)views_.size() <= stream_id) views_.resize(stream_id + 1);
        if (!views_[stream_id] || views_[stream_id]->stream() != s) views_[stream_id].reset(new device_stream(s));
        return *views_[stream_id];
    }
    device_event enqueue_event(int stream_id) const { return device_stream(get_stream(stream_id)).enqueue_event(); }
    // the stream finished C blocks leave the device on (gpu_context.cpp:55-57)
    device_stream& get_result_stream() {
        if (!result_) result_.reset(new device_stream(static_cast<runtime_api::StreamType>(checked(TMM_STREAM_D2H, 0))));
        return *result_;
    }

    int get_num_streams() const { return tmm_context_get_num_streams(ctx_); }
    void set_num_streams(int streams) {
        int tm = 0, tn = 0, tk = 0;
        check_tmm_status(tmm_context_get_max_tile_sizes(ctx_, &tm, &tn, &tk));
        check_tmm_status(tmm_context_set_streams_and_tiles(ctx_, streams, tm, tn, tk));
    }

    tmm_context* native() const { return ctx_; }

private:
    void* checked(int kind, int index) const {
        void* s = tmm_context_stream(ctx_, kind, index);
        if (!s) throw std::runtime_error("stream id has to be in the range [0, n_streams)");
        return s;
    }
    tmm_context* ctx_;
    std::vector<std::unique_ptr<device_stream>> views_;  // non-owning views of the context's streams, made on demand
    std::unique_ptr<device_stream> result_;
};

}  // namespace gpu
