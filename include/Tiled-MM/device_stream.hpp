// Drop-in counterpart of reference src/Tiled-MM/device_stream.hpp: move-only owner of one non-blocking CUDA stream
// with enqueue_event / wait_on_event.  It can also be a non-owning view of a stream that belongs to a context
// (gpu_context::get_device_stream), which the reference has no need for because its gpu_context owns device_streams.
#pragma once
#include "device_event.hpp"
#include "gpu_runtime_api.hpp"
#include "util.hpp"

#include <utility>

namespace gpu {

class device_stream {
public:
    device_stream() {
        check_runtime_status(runtime_api::stream_create_with_flags(&stream_, runtime_api::flag::StreamNonBlocking));
        owns_ = true;
    }
    // view of an existing stream: never destroyed here
    explicit device_stream(runtime_api::StreamType borrowed) : stream_(borrowed), owns_(false) {}
    ~device_stream() { reset(); }

    device_stream(device_stream&& other) noexcept : stream_(other.stream_), owns_(std::exchange(other.owns_, false)) {}
    device_stream& operator=(device_stream&& other) noexcept {
        if (this != &other) {
            reset();
            stream_ = other.stream_;
            owns_ = std::exchange(other.owns_, false);
        }
        return *this;
    }
    device_stream(device_stream&) = delete;
    device_stream& operator=(device_stream&) = delete;

    runtime_api::StreamType stream() const { return stream_; }

    // a fresh event recorded behind everything queued on this stream so far
    device_event enqueue_event() const {
        device_event e;
        check_runtime_status(runtime_api::event_record(e.get(), stream_));
        return e;
    }

    // later work on this stream starts only after `e` has happened
    void wait_on_event(device_event& e) const { check_runtime_status(runtime_api::stream_wait_event(stream_, e.get(), 0)); }

private:
    void reset() {
        if (owns_) runtime_api::stream_destroy(stream_);
        owns_ = false;
    }
    runtime_api::StreamType stream_{};
    bool owns_ = false;
};

}  // namespace gpu
