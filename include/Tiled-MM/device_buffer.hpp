// Drop-in counterpart of reference src/Tiled-MM/device_buffer.hpp: n_streams tile-sized slabs in one device_vector,
// slab s starting at s * tile.size().  The scheduler of this library does not stage through these (it owns panel / ring
// storage sized per call, csrc/tmm_context.cu); mm_handle::get_device_buffer_{a,b,c} hand them out for callers that use
// the handle as their per-stream device scratch, and they are only allocated once somebody asks for a pointer.
// Differences kept on purpose: a bad stream id throws (the reference builds the exception and drops it,
// device_buffer.hpp:46-48); offsets are 64-bit (Q1).
#pragma once
#include "device_vector.hpp"
#include "tile_dim.hpp"

#include <stdexcept>

namespace gpu {

template <typename T>
class device_buffer {
public:
    device_buffer() = default;
    explicit device_buffer(int streams) : n_streams_(streams) {}

    T* stream_buffer(int stream_id) {
        if (stream_id < 0 || stream_id >= n_streams_) throw std::runtime_error("stream id in device buffer has to be in the range [0, n_streams)");
        materialise();
        return d_vec_.data() + (std::size_t)stream_id * tile_.size64();
    }

    T* data() {
        materialise();
        return d_vec_.data();
    }

    void set_num_streams(int streams) { n_streams_ = streams; dirty_ = true; }
    void set_tile_sizes(tile_dim tile) { tile_ = tile; dirty_ = true; }
    tile_dim get_tile_sizes() { return tile_; }
    void set_streams_and_tiles(int streams, tile_dim tile) { n_streams_ = streams; tile_ = tile; dirty_ = true; }

    int get_num_streams() const { return n_streams_; }
    // elements the slabs span: n_streams * tile rows * tile cols
    std::size_t size() const { return (std::size_t)(n_streams_ > 0 ? n_streams_ : 0) * tile_.size64(); }
    bool allocated() { return d_vec_.capacity() > 0; }

private:
    void materialise() {
        if (dirty_) { d_vec_.resize(size()); dirty_ = false; }  // grow-only, contents not preserved (device_vector.hpp:92-107)
    }
    int n_streams_ = 0;
    tile_dim tile_;
    device_vector<T> d_vec_;
    bool dirty_ = true;
};

}  // namespace gpu
