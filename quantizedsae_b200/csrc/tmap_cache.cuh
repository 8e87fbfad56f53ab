// Per-process cache of encoded TMA descriptors ("no global state except a per-device TMA-descriptor
// cache"). A CUtensorMap depends only on (device, base pointer, shape, pitch, box, data type, swizzle) -- not on the
// data -- and cuTensorMapEncodeTiled costs ~1 us of host time; a direct (non-graph) forward encodes four of them. The
// steady state of a serving loop presents the same few (pointer, shape) keys again and again (prepared weights,
// per-stream workspaces, rotating inputs): 256 direct-mapped entries, one mutex, overwritten on collision.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

namespace qsae {

struct TmapKey {
  const void* base;
  unsigned long long rows, cols, pitch_bytes;
  unsigned int box_cols, box_rows;
  int dtype, swizzle, device;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && pitch_bytes == o.pitch_bytes && box_cols == o.box_cols &&
           box_rows == o.box_rows && dtype == o.dtype && swizzle == o.swizzle && device == o.device;
  }
};

class TmapCache {
 public:
  // encode(map) is called on a miss and must return true on success
  template <typename Encode>
  bool get(const TmapKey& key, CUtensorMap* out, Encode encode) {
    const size_t slot = hash(key) % kSlots;
    {
      std::lock_guard<std::mutex> g(mu_);
      if (valid_[slot] && keys_[slot] == key) {
        memcpy(out, &maps_[slot], sizeof(CUtensorMap));
        ++hits_;
        return true;
      }
    }
    if (!encode(out)) return false;
    std::lock_guard<std::mutex> g(mu_);
    keys_[slot] = key;
    memcpy(&maps_[slot], out, sizeof(CUtensorMap));
    valid_[slot] = true;
    ++misses_;
    return true;
  }
  unsigned long long hits() const { return hits_; }
  unsigned long long misses() const { return misses_; }

 private:
  static constexpr size_t kSlots = 256;
  static size_t hash(const TmapKey& k) {
    unsigned long long h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows * 0xC2B2AE3D27D4EB4Full) ^ (k.cols << 17) ^ (static_cast<unsigned long long>(k.box_rows) << 40) ^
         (static_cast<unsigned long long>(k.box_cols) << 52) ^ (static_cast<unsigned long long>(k.device) << 60) ^ k.pitch_bytes;
    return static_cast<size_t>(h ^ (h >> 29));
  }
  std::mutex mu_;
  TmapKey keys_[kSlots];
  CUtensorMap maps_[kSlots];
  bool valid_[kSlots] = {false};
  unsigned long long hits_ = 0, misses_ = 0;
};

// one instance for the whole library (defined in encode_topk_sm100.cu)
TmapCache& tmap_cache();
inline int tmap_current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

}  // namespace qsae
