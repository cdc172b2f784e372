// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT tsl::robin_map.
//
// Stand-in for Tessil's tsl::robin_map v1.0.1 (cpp/sage_icp/3rdparty/tsl_robin/tsl_robin.cmake:24; not in this image), limited
// to what the reference's hot-path sources call: default/copy construction, reserve, find, end, contains, insert(pair),
// erase(key), clear, empty, size, iteration.  Written from the published rules of that version — power-of-two growth by 2 from
// 0 buckets, max_load_factor 0.5 with load_threshold = size_t(float(bucket_count) * 0.5f), robin-hood insertion with
// swap-and-carry, rehash by re-inserting the old bucket array in order, backward-shift deletion, no shrinking (min load factor
// 0), growth forced by a probe length above 8192 — because the reference OBSERVES the resulting iteration order
// (core/Preprocessing.cpp:78, core/VoxelHashMap.cpp:135) and erases while iterating (core/VoxelHashMap.cpp:177-183).
// It is a generic-key sibling of oracle/robin_table.hpp; both are pinned against the pure-Python model in
// tests/test_oracle_robin.py, neither against the real tsl source.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <type_traits>
#include <utility>
#include <vector>

namespace tsl {

template <class Key, class T, class Hash = std::hash<Key>, class KeyEqual = std::equal_to<Key>>
class robin_map {
    struct Bucket {
        int dist = -1;  // distance from the ideal bucket; -1 = empty
        std::size_t hash = 0;
        std::pair<Key, T> kv;
    };
    static constexpr int kDistLimit = 8192;

public:
    using value_type = std::pair<Key, T>;

    template <bool Const>
    class iter {
        using map_ptr = typename std::conditional<Const, const robin_map *, robin_map *>::type;
        map_ptr m_ = nullptr;
        std::size_t i_ = 0;
        friend class robin_map;
        template <bool>
        friend class iter;
        void skip() {
            while (i_ < m_->b_.size() && m_->b_[i_].dist < 0) ++i_;
        }

    public:
        iter() = default;
        iter(map_ptr m, std::size_t i) : m_(m), i_(i) {}
        iter(const iter<false> &o) : m_(o.m_), i_(o.i_) {}
        const value_type &operator*() const { return m_->b_[i_].kv; }
        const value_type *operator->() const { return &m_->b_[i_].kv; }
        const Key &key() const { return m_->b_[i_].kv.first; }
        typename std::conditional<Const, const T &, T &>::type value() const { return m_->b_[i_].kv.second; }
        iter &operator++() {
            ++i_;
            skip();
            return *this;
        }
        bool operator==(const iter &o) const { return i_ == o.i_; }
        bool operator!=(const iter &o) const { return i_ != o.i_; }
    };
    using iterator = iter<false>;
    using const_iterator = iter<true>;

    robin_map() = default;

    iterator begin() {
        iterator it(this, 0);
        it.skip();
        return it;
    }
    iterator end() { return iterator(this, b_.size()); }
    const_iterator begin() const {
        const_iterator it(this, 0);
        it.skip();
        return it;
    }
    const_iterator end() const { return const_iterator(this, b_.size()); }

    bool empty() const { return n_ == 0; }
    std::size_t size() const { return n_; }
    std::size_t bucket_count() const { return b_.size(); }
    void clear() {  // keeps the bucket count
        for (auto &b : b_) b = Bucket{};
        n_ = 0;
        grow_next_ = false;
    }
    void reserve(std::size_t count) { rehash_to((std::size_t)std::ceil((float)count / 0.5f)); }

    iterator find(const Key &k) { return iterator(this, locate(k)); }
    const_iterator find(const Key &k) const { return const_iterator(this, locate(k)); }
    bool contains(const Key &k) const { return locate(k) != b_.size(); }

    std::pair<iterator, bool> insert(const value_type &v) {
        const std::size_t h = Hash{}(v.first);
        std::size_t i = 0;
        int d = 0;
        if (!b_.empty()) {
            const std::size_t mask = b_.size() - 1;
            for (i = h & mask; d <= b_[i].dist; ++d, i = (i + 1) & mask)
                if (KeyEqual{}(b_[i].kv.first, v.first)) return {iterator(this, i), false};
        }
        while (grow_next_ || d > kDistLimit || n_ >= threshold_) {
            grow(b_.empty() ? 2 : b_.size() * 2);
            grow_next_ = false;
            const std::size_t mask = b_.size() - 1;
            for (i = h & mask, d = 0; d <= b_[i].dist; ++d) i = (i + 1) & mask;
        }
        const std::size_t mask = b_.size() - 1, home = i;
        Bucket e;
        e.dist = d, e.hash = h, e.kv = v;
        if (b_[i].dist >= 0) {
            std::swap(e, b_[i]);
            for (d = e.dist + 1, i = (i + 1) & mask; b_[i].dist >= 0; ++d, i = (i + 1) & mask) {
                if (d <= b_[i].dist) continue;
                if (d >= kDistLimit) grow_next_ = true;
                e.dist = d;
                std::swap(e, b_[i]);
                d = e.dist;
            }
            e.dist = d;
        }
        b_[i] = std::move(e);
        ++n_;
        return {iterator(this, home), true};
    }

    std::size_t erase(const Key &k) {
        std::size_t i = locate(k);
        if (i == b_.size()) return 0;
        const std::size_t mask = b_.size() - 1;
        b_[i] = Bucket{};
        --n_;
        for (std::size_t next = (i + 1) & mask; b_[next].dist > 0; i = next, next = (next + 1) & mask) {  // backward shift
            b_[i] = std::move(b_[next]);
            b_[i].dist -= 1;
            b_[next] = Bucket{};
        }
        return 1;
    }

private:
    std::size_t locate(const Key &k) const {
        if (b_.empty()) return 0;  // == b_.size()
        const std::size_t mask = b_.size() - 1;
        std::size_t i = Hash{}(k) & mask;
        for (int d = 0; d <= b_[i].dist; ++d, i = (i + 1) & mask)
            if (KeyEqual{}(b_[i].kv.first, k)) return i;
        return b_.size();
    }
    void rehash_to(std::size_t count) {  // rehash(count): at least what the load factor needs, rounded up to a power of two
        const std::size_t need = (std::size_t)std::ceil((float)n_ / 0.5f);
        if (count < need) count = need;
        std::size_t pow2 = 1;
        while (pow2 < count) pow2 *= 2;
        if (count == 0) pow2 = 0;
        if (pow2 != b_.size()) grow(pow2);
    }
    void grow(std::size_t count) {
        std::vector<Bucket> old;
        old.swap(b_);
        b_.assign(count, Bucket{});
        threshold_ = (std::size_t)((float)count * 0.5f);
        if (count == 0) return;
        const std::size_t mask = count - 1;
        for (auto &o : old) {
            if (o.dist < 0) continue;
            Bucket e = std::move(o);
            std::size_t i = e.hash & mask;
            for (int d = 0;; ++d, i = (i + 1) & mask) {
                if (d <= b_[i].dist) continue;
                e.dist = d;
                if (b_[i].dist < 0) {
                    b_[i] = std::move(e);
                    break;
                }
                std::swap(e, b_[i]);
                d = e.dist;
            }
        }
    }

    std::vector<Bucket> b_;
    std::size_t n_ = 0, threshold_ = 0;
    bool grow_next_ = false;
};

}  // namespace tsl
