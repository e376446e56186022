// Drop-in counterpart of reference src/Tiled-MM/device_event.hpp: move-only owner of one CUDA event.
// Same members (get, wait, time_since).  Two differences, both supersets: the event is created WITH timing when asked
// (the reference creates every event with EventDisableTiming, device_event.hpp:12-16, so its own time_since can only fail),
// and a moved-from object is inert.  The scheduler does not use this class: it draws events from a per-context pool
// (csrc/tmm_internal.h, tmm_context::get_event) instead of creating one per tile.
#pragma once
#include "gpu_runtime_api.hpp"
#include "util.hpp"

#include <utility>

namespace gpu {

class device_event {
public:
    explicit device_event(bool with_timing = false) {
        check_runtime_status(runtime_api::event_create_with_flags(&event_, with_timing ? 0u : (unsigned)runtime_api::flag::EventDisableTiming));
        owns_ = true;
    }
    ~device_event() { reset(); }

    device_event(device_event&& other) noexcept : event_(other.event_), owns_(std::exchange(other.owns_, false)) {}
    device_event& operator=(device_event&& other) noexcept {
        if (this != &other) {
            reset();
            event_ = other.event_;
            owns_ = std::exchange(other.owns_, false);
        }
        return *this;
    }
    device_event(device_event&) = delete;
    device_event& operator=(device_event&) = delete;

    runtime_api::EventType& get() { return event_; }

    // block the host until the event has happened
    void wait() { check_runtime_status(runtime_api::event_synchronize(event_)); }

    // seconds between `other` (earlier) and this event; both need timing enabled
    double time_since(device_event& other) {
        float ms = 0.0f;
        check_runtime_status(runtime_api::event_elapsed_time(&ms, other.get(), event_));
        return double(ms) / 1.e3;
    }

private:
    void reset() {
        if (owns_) runtime_api::event_destroy(event_);
        owns_ = false;
    }
    runtime_api::EventType event_{};
    bool owns_ = false;
};

}  // namespace gpu
