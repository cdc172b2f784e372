// ORACLE — TEST INFRASTRUCTURE ONLY (see se3.hpp header note).
//
// Sequential restatement of tsl::robin_map v1.0.1 (Tessil/robin-map, pinned at
// cpp/sage_icp/3rdparty/tsl_robin/tsl_robin.cmake:24; source NOT in /root/reference), limited to what the
// reference observes: find / insert / erase / clear and **iteration order**, which is consumed at
//   core/Preprocessing.cpp:78   (VoxelDownsample output order)
//   core/VoxelHashMap.cpp:135   (Pointcloud() order)
//   core/VoxelHashMap.cpp:177   (RemovePointsFarFromLocation erase-while-iterating)
// Published algorithm restated here:
//   * power-of-two growth policy (factor 2), initial bucket count 0, max_load_factor 0.5,
//     load_threshold = size_t(float(bucket_count) * 0.5f);
//   * lookup: walk from ideal bucket while dist <= resident.dist;
//   * insert (after a failed lookup): if size() >= load_threshold (or grow flag / probe limit) rehash to
//     next_bucket_count() = (mask + 1) * 2 and redo the walk; place into an empty bucket or robin-hood
//     swap-and-carry (carried element continues with the evicted resident's distance);
//   * rehash: iterate old bucket array in order, re-insert each with the same rule;
//   * erase: clear bucket, backward-shift following buckets while their dist > 0; never shrinks.
//   * DIST_FROM_IDEAL_BUCKET_LIMIT = 8192 sets grow_on_next_insert (only reachable when the 20-bit hash
//     saturates).
// PARITY UNPINNED against the real tsl source (absent, no network); pinned instead against an independent
// Python emulator of the same published rules in tests/test_oracle_robin.py.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

namespace orc {

struct Voxel {
    int32_t x, y, z;
    bool operator==(const Voxel &o) const { return x == o.x && y == o.y && z == o.z; }
};

// VoxelHash — core/VoxelHashMap.hpp:72-77 and core/Preprocessing.cpp:35-40 (u32 wrap, 20-bit mask)
inline uint32_t voxel_hash(const Voxel &v) {
    return ((1u << 20) - 1u) &
           ((uint32_t)v.x * 73856093u ^ (uint32_t)v.y * 19349663u ^ (uint32_t)v.z * 83492791u);
}

template <class T>
class RobinTable {
public:
    static constexpr int16_t kEmpty = -1;
    static constexpr int kDistLimit = 8192;
    struct Bucket {
        int16_t dist = kEmpty;
        Voxel key{0, 0, 0};
        T value{};
        bool empty() const { return dist == kEmpty; }
    };

    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    size_t bucket_count() const { return buckets_.size(); }
    void clear() {  // tsl clear(): empties buckets, keeps bucket_count
        for (auto &b : buckets_) b = Bucket{};
        n_ = 0;
        grow_next_ = false;
    }
    // iteration: bucket array order, skipping empties
    const std::vector<Bucket> &buckets() const { return buckets_; }
    std::vector<Bucket> &buckets() { return buckets_; }

    // returns bucket index or npos
    static constexpr size_t npos = (size_t)-1;
    size_t find(const Voxel &k) const {
        if (buckets_.empty()) return npos;
        size_t ib = voxel_hash(k) & mask_;
        int d = 0;
        while (d <= buckets_[ib].dist) {
            if (buckets_[ib].key == k) return ib;
            ib = (ib + 1) & mask_;
            ++d;
        }
        return npos;
    }
    bool contains(const Voxel &k) const { return find(k) != npos; }
    T &value_at(size_t ib) { return buckets_[ib].value; }
    const T &value_at(size_t ib) const { return buckets_[ib].value; }

    // insert(key,value); no-op if present. Returns true if inserted.
    bool insert(const Voxel &k, T v) {
        const uint32_t h = voxel_hash(k);
        size_t ib = buckets_.empty() ? 0 : (h & mask_);
        int d = 0;
        if (!buckets_.empty()) {
            while (d <= buckets_[ib].dist) {
                if (buckets_[ib].key == k) return false;
                ib = (ib + 1) & mask_;
                ++d;
            }
        }
        while (rehash_on_extreme_load(d)) {
            ib = h & mask_;
            d = 0;
            while (d <= buckets_[ib].dist) {
                ib = (ib + 1) & mask_;
                ++d;
            }
        }
        Bucket carry;
        carry.dist = (int16_t)d;
        carry.key = k;
        carry.value = std::move(v);
        if (buckets_[ib].empty()) {
            buckets_[ib] = std::move(carry);
        } else {
            std::swap(carry, buckets_[ib]);
            ib = (ib + 1) & mask_;
            int cd = carry.dist + 1;
            while (!buckets_[ib].empty()) {
                if (cd > buckets_[ib].dist) {
                    if (cd >= kDistLimit) grow_next_ = true;
                    carry.dist = (int16_t)cd;
                    std::swap(carry, buckets_[ib]);
                    cd = carry.dist;
                }
                ib = (ib + 1) & mask_;
                ++cd;
            }
            carry.dist = (int16_t)cd;
            buckets_[ib] = std::move(carry);
        }
        ++n_;
        return true;
    }

    // erase by bucket index with backward-shift deletion
    void erase_at(size_t ib) {
        buckets_[ib] = Bucket{};
        --n_;
        size_t prev = ib, cur = (ib + 1) & mask_;
        while (buckets_[cur].dist > 0) {
            buckets_[prev] = std::move(buckets_[cur]);
            buckets_[prev].dist = (int16_t)(buckets_[prev].dist - 1);
            buckets_[cur] = Bucket{};
            prev = cur;
            cur = (cur + 1) & mask_;
        }
    }

private:
    bool rehash_on_extreme_load(int cur_dist) {
        if (grow_next_ || cur_dist > kDistLimit || n_ >= load_threshold_) {
            rehash(buckets_.empty() ? 2 : buckets_.size() * 2);
            grow_next_ = false;
            return true;
        }
        return false;
    }
    void rehash(size_t count) {
        std::vector<Bucket> old;
        old.swap(buckets_);
        buckets_.assign(count, Bucket{});
        mask_ = count - 1;
        load_threshold_ = (size_t)((float)count * 0.5f);
        for (auto &b : old) {
            if (b.empty()) continue;
            size_t ib = voxel_hash(b.key) & mask_;
            Bucket carry = std::move(b);
            int d = 0;
            while (true) {  // insert_value_on_rehash
                if (d > buckets_[ib].dist) {
                    if (buckets_[ib].empty()) {
                        carry.dist = (int16_t)d;
                        buckets_[ib] = std::move(carry);
                        break;
                    }
                    carry.dist = (int16_t)d;
                    std::swap(carry, buckets_[ib]);
                    d = carry.dist;
                }
                ++d;
                ib = (ib + 1) & mask_;
            }
        }
    }

    std::vector<Bucket> buckets_;
    size_t mask_ = 0;
    size_t n_ = 0;
    size_t load_threshold_ = 0;
    bool grow_next_ = false;
};

}  // namespace orc
