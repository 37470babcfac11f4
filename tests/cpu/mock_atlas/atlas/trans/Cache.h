// Mock of trans/Cache.h:41-136: a Cache hands out raw byte entries; only the Legendre entry matters to a backend.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>
namespace atlas {
namespace trans {
class TransCacheEntry {
public:
    TransCacheEntry() = default;
    TransCacheEntry(const void* data, size_t size): data_(data), size_(size) {}
    operator bool() const { return size_ != 0; }
    size_t size() const { return size_; }
    const void* data() const { return data_; }
private:
    const void* data_ = nullptr;
    size_t size_ = 0;
};
class Cache {
public:
    Cache() = default;
    explicit Cache(const TransCacheEntry& legendre): legendre_(legendre) {}
    const TransCacheEntry& legendre() const { return legendre_; }
protected:
    TransCacheEntry legendre_;
    std::shared_ptr<std::vector<char>> store_;   // entries are shared_ptr in the reference (Cache.h:113-117): copies of a Cache
                                                 // (also sliced from a LegendreCache) keep the memory alive
};
class LegendreCache : public Cache {  // Cache.h:123-128: LegendreCache(size) owns its (host) memory
public:
    explicit LegendreCache(size_t size) {
        store_ = std::make_shared<std::vector<char>>(size);
        legendre_ = TransCacheEntry(store_->data(), size);
    }
    LegendreCache(const void* address, size_t size) { legendre_ = TransCacheEntry(address, size); }
};
}  // namespace trans
}  // namespace atlas
