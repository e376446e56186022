// Drop-in counterpart of reference src/Tiled-MM/device_vector.hpp: grow-only device array with 1.2x slack
// whose resize() does NOT preserve contents (device_vector.hpp:92-107).  Additionally it can be a
// non-owning view: mm_handle::get_full_device_buffer_c() returns one that tracks the context's device C.
#pragma once
#include "util.hpp"

#include <cmath>
#include <cstddef>

namespace gpu {

template <typename T>
class device_vector {
public:
    device_vector() = default;
    explicit device_vector(std::size_t n) : size_(n), capacity_((std::size_t)std::ceil(1.2 * n)) { data_ = malloc_device<T>(capacity_); }

    device_vector(device_vector& other) = delete;
    device_vector& operator=(device_vector&) = delete;

    device_vector& operator=(device_vector&& other) {
        if (this != &other) {
            release();
            data_ = other.data_; size_ = other.size_; capacity_ = other.capacity_; view_of_ = other.view_of_;
            other.data_ = nullptr; other.size_ = 0; other.capacity_ = 0; other.view_of_ = nullptr;
        }
        return *this;
    }

    T* data() { return view_of_ ? static_cast<T*>(tmm_context_device_c(view_of_)) : data_; }
    std::size_t size() { return view_of_ ? tmm_context_device_c_size(view_of_) : size_; }
    std::size_t capacity() { return view_of_ ? size() : capacity_; }

    void resize(std::size_t size) {
        if (view_of_ || size == 0) return;  // a view is sized by the context (set_full_sizes / gemm)
        if (size > capacity_) {
            release();
            size_ = size;
            capacity_ = (std::size_t)std::ceil(1.2 * size);
            data_ = malloc_device<T>(capacity_);
        } else {
            size_ = size;
        }
    }

    // library-internal: make this object a window onto a context's full device C
    void bind_to_context(tmm_context* ctx) { release(); view_of_ = ctx; }

    ~device_vector() { release(); }

private:
    void release() {
        if (!view_of_ && capacity_ > 0 && data_) tmm_free_device(data_);
        data_ = nullptr; size_ = 0; capacity_ = 0;
    }
    T* data_ = nullptr;
    std::size_t size_ = 0;
    std::size_t capacity_ = 0;
    tmm_context* view_of_ = nullptr;
};

}  // namespace gpu
