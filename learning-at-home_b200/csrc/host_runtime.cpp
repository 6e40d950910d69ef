// host_runtime.cpp — native host-side runtime pieces (plain C ABI, loaded with ctypes; no Python.h, no torch).
//
//  * batch assembly:   multi-threaded gather/scatter of row blocks between task tensors and one (pinned) staging buffer
//                      (the reference does np.concatenate into shared memory: lib/runtime/task_pool.py:162-168)
//  * wire framing:     blocking send-all / recv-exact loops and the 4+8 byte message header of the TCP fallback path
//                      (the reference loops over 2 KiB recv() calls in Python: lib/utils/connection.py:36-48)
//  * DHT routing:      Kademlia k-bucket routing table with XOR metric on 160-bit ids (the reference delegates to the
//                      third-party `kademlia` package: lib/network/__init__.py:6,20)
//  * expert index:     open-addressing hash table uid-hash -> (owner, slot, heartbeat) used by the in-box network facade
//
// All calls are made with the GIL released (ctypes does that), so handler threads really run in parallel here.
#include <algorithm>
#include <array>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <sys/socket.h>
#include <sys/types.h>
#include <unistd.h>

extern "C" {

// ------------------------------------------------------------------------------------------------ batch assembly
// dst = concat(srcs[i][:nbytes[i]]); large copies are split over `threads` worker threads
int lah_host_gather(const void** srcs, const unsigned long long* nbytes, int n, void* dst, int threads) {
    std::vector<unsigned long long> off(n + 1, 0);
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + nbytes[i];
    const unsigned long long total = off[n];
    if (threads <= 1 || total < (1ull << 20)) {
        for (int i = 0; i < n; ++i) memcpy(static_cast<char*>(dst) + off[i], srcs[i], nbytes[i]);
        return 0;
    }
    threads = std::min(threads, 16);
    const unsigned long long chunk = (total + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            const unsigned long long lo = t * chunk, hi = std::min(total, (t + 1) * chunk);
            for (int i = 0; i < n; ++i) {
                const unsigned long long a = std::max(lo, off[i]), b = std::min(hi, off[i + 1]);
                if (a < b) memcpy(static_cast<char*>(dst) + a, static_cast<const char*>(srcs[i]) + (a - off[i]), b - a);
            }
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}

// inverse: dsts[i][:nbytes[i]] = src[off_i : off_i + nbytes[i]]
int lah_host_scatter(const void* src, void** dsts, const unsigned long long* nbytes, int n) {
    unsigned long long off = 0;
    for (int i = 0; i < n; ++i) {
        memcpy(dsts[i], static_cast<const char*>(src) + off, nbytes[i]);
        off += nbytes[i];
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ wire framing
long long lah_host_send_all(int fd, const void* buf, unsigned long long n) {
    const char* p = static_cast<const char*>(buf);
    unsigned long long sent = 0;
    while (sent < n) {
        const ssize_t r = ::send(fd, p + sent, n - sent, MSG_NOSIGNAL);
        if (r < 0) {
            if (errno == EINTR) continue;
            return -errno;
        }
        if (r == 0) return -EPIPE;
        sent += static_cast<unsigned long long>(r);
    }
    return static_cast<long long>(sent);
}

long long lah_host_recv_exact(int fd, void* buf, unsigned long long n) {
    char* p = static_cast<char*>(buf);
    unsigned long long got = 0;
    while (got < n) {
        const ssize_t r = ::recv(fd, p + got, n - got, 0);
        if (r < 0) {
            if (errno == EINTR) continue;
            return -errno;
        }
        if (r == 0) return -ECONNRESET;  // peer closed
        got += static_cast<unsigned long long>(r);
    }
    return static_cast<long long>(got);
}

// header(4 ASCII) + length(8, big-endian) + payload
int lah_host_send_message(int fd, const char* header4, const void* payload, unsigned long long n) {
    unsigned char prefix[12];
    memcpy(prefix, header4, 4);
    for (int i = 0; i < 8; ++i) prefix[4 + i] = static_cast<unsigned char>((n >> (8 * (7 - i))) & 0xFF);
    long long r = lah_host_send_all(fd, prefix, 12);
    if (r < 0) return static_cast<int>(r);
    r = lah_host_send_all(fd, payload, n);
    return r < 0 ? static_cast<int>(r) : 0;
}

int lah_host_recv_header(int fd, char* header4_out, unsigned long long* len_out) {
    unsigned char prefix[12];
    const long long r = lah_host_recv_exact(fd, prefix, 12);
    if (r < 0) return static_cast<int>(r);
    memcpy(header4_out, prefix, 4);
    unsigned long long n = 0;
    for (int i = 0; i < 8; ++i) n = (n << 8) | prefix[4 + i];
    *len_out = n;
    return 0;
}

// ------------------------------------------------------------------------------------------------ Kademlia routing table
namespace {
constexpr int ID_BYTES = 20;
using NodeId = std::array<uint8_t, ID_BYTES>;

struct Contact {
    NodeId id;
    uint32_t ip;
    uint16_t port;
    double last_seen;
};

struct RoutingTable {
    NodeId self;
    int k;
    std::vector<std::vector<Contact>> buckets;  // bucket b holds ids whose XOR distance has its top set bit at b
    std::mutex mu;
};

inline int bucket_index(const NodeId& a, const NodeId& b) {
    for (int i = 0; i < ID_BYTES; ++i) {
        const uint8_t x = a[i] ^ b[i];
        if (x) {
            int bit = 7;
            while (!((x >> bit) & 1)) --bit;
            return (ID_BYTES - 1 - i) * 8 + bit;
        }
    }
    return -1;  // identical ids
}

inline bool closer(const NodeId& target, const NodeId& a, const NodeId& b) {
    for (int i = 0; i < ID_BYTES; ++i) {
        const uint8_t da = a[i] ^ target[i], db = b[i] ^ target[i];
        if (da != db) return da < db;
    }
    return false;
}
}  // namespace

void* lah_rt_create(const uint8_t* self_id, int k) {
    auto* rt = new RoutingTable();
    memcpy(rt->self.data(), self_id, ID_BYTES);
    rt->k = k;
    rt->buckets.resize(ID_BYTES * 8);
    return rt;
}

void lah_rt_destroy(void* h) { delete static_cast<RoutingTable*>(h); }

// returns 1: inserted / refreshed; 0: bucket full (least-recently-seen contact is written to evict_* for a liveness ping);
// -1: own id
int lah_rt_add(void* h, const uint8_t* id, uint32_t ip, uint16_t port, double now, uint8_t* evict_id, uint32_t* evict_ip,
               uint16_t* evict_port) {
    auto* rt = static_cast<RoutingTable*>(h);
    NodeId nid;
    memcpy(nid.data(), id, ID_BYTES);
    const int b = bucket_index(rt->self, nid);
    if (b < 0) return -1;
    std::lock_guard<std::mutex> lock(rt->mu);
    auto& bucket = rt->buckets[b];
    for (size_t i = 0; i < bucket.size(); ++i) {
        if (bucket[i].id == nid) {  // move to the tail (most recently seen)
            Contact c = bucket[i];
            c.ip = ip; c.port = port; c.last_seen = now;
            bucket.erase(bucket.begin() + i);
            bucket.push_back(c);
            return 1;
        }
    }
    if (static_cast<int>(bucket.size()) < rt->k) {
        bucket.push_back(Contact{nid, ip, port, now});
        return 1;
    }
    if (evict_id) {
        memcpy(evict_id, bucket.front().id.data(), ID_BYTES);
        *evict_ip = bucket.front().ip;
        *evict_port = bucket.front().port;
    }
    return 0;
}

int lah_rt_remove(void* h, const uint8_t* id) {
    auto* rt = static_cast<RoutingTable*>(h);
    NodeId nid;
    memcpy(nid.data(), id, ID_BYTES);
    const int b = bucket_index(rt->self, nid);
    if (b < 0) return 0;
    std::lock_guard<std::mutex> lock(rt->mu);
    auto& bucket = rt->buckets[b];
    for (size_t i = 0; i < bucket.size(); ++i)
        if (bucket[i].id == nid) {
            bucket.erase(bucket.begin() + i);
            return 1;
        }
    return 0;
}

int lah_rt_size(void* h) {
    auto* rt = static_cast<RoutingTable*>(h);
    std::lock_guard<std::mutex> lock(rt->mu);
    int n = 0;
    for (auto& b : rt->buckets) n += static_cast<int>(b.size());
    return n;
}

// the n contacts closest (XOR metric) to target; returns how many were written
int lah_rt_closest(void* h, const uint8_t* target, int n, uint8_t* out_ids, uint32_t* out_ips, uint16_t* out_ports) {
    auto* rt = static_cast<RoutingTable*>(h);
    NodeId t;
    memcpy(t.data(), target, ID_BYTES);
    std::vector<Contact> all;
    {
        std::lock_guard<std::mutex> lock(rt->mu);
        for (auto& b : rt->buckets) all.insert(all.end(), b.begin(), b.end());
    }
    const int m = std::min<int>(n, static_cast<int>(all.size()));
    std::partial_sort(all.begin(), all.begin() + m, all.end(),
                      [&](const Contact& a, const Contact& b) { return closer(t, a.id, b.id); });
    for (int i = 0; i < m; ++i) {
        memcpy(out_ids + i * ID_BYTES, all[i].id.data(), ID_BYTES);
        out_ips[i] = all[i].ip;
        out_ports[i] = all[i].port;
    }
    return m;
}

// ------------------------------------------------------------------------------------------------ in-box expert index
// open addressing table: 64-bit uid hash -> (owner rank, local slot, heartbeat time)
namespace {
struct IndexEntry {
    uint64_t key;  // 0 = empty
    int32_t owner, slot;
    double heartbeat;
};
struct ExpertIndex {
    std::vector<IndexEntry> table;
    std::mutex mu;
    size_t used = 0;
};
inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x ? x : 1;
}
}  // namespace

unsigned long long lah_hash_bytes(const uint8_t* data, int n) {
    uint64_t h = 1469598103934665603ULL;  // FNV-1a, then avalanche
    for (int i = 0; i < n; ++i) {
        h ^= data[i];
        h *= 1099511628211ULL;
    }
    return mix64(h);
}

void* lah_index_create(int capacity_pow2) {
    auto* ix = new ExpertIndex();
    size_t cap = 16;
    while (cap < static_cast<size_t>(capacity_pow2)) cap <<= 1;
    ix->table.assign(cap, IndexEntry{0, -1, -1, 0.0});
    return ix;
}

void lah_index_destroy(void* h) { delete static_cast<ExpertIndex*>(h); }

int lah_index_put(void* h, unsigned long long key, int owner, int slot, double heartbeat) {
    auto* ix = static_cast<ExpertIndex*>(h);
    std::lock_guard<std::mutex> lock(ix->mu);
    if ((ix->used + 1) * 2 > ix->table.size()) {  // grow + rehash
        std::vector<IndexEntry> old;
        old.swap(ix->table);
        ix->table.assign(old.size() * 2, IndexEntry{0, -1, -1, 0.0});
        ix->used = 0;
        for (auto& e : old)
            if (e.key) {
                size_t i = e.key & (ix->table.size() - 1);
                while (ix->table[i].key) i = (i + 1) & (ix->table.size() - 1);
                ix->table[i] = e;
                ++ix->used;
            }
    }
    size_t i = key & (ix->table.size() - 1);
    while (ix->table[i].key && ix->table[i].key != key) i = (i + 1) & (ix->table.size() - 1);
    if (!ix->table[i].key) ++ix->used;
    ix->table[i] = IndexEntry{key, owner, slot, heartbeat};
    return 0;
}

// returns 1 if found (and fresh enough when max_age >= 0), else 0
int lah_index_get(void* h, unsigned long long key, double now, double max_age, int* owner, int* slot, double* heartbeat) {
    auto* ix = static_cast<ExpertIndex*>(h);
    std::lock_guard<std::mutex> lock(ix->mu);
    size_t i = key & (ix->table.size() - 1);
    while (ix->table[i].key) {
        if (ix->table[i].key == key) {
            if (owner) *owner = ix->table[i].owner;
            if (slot) *slot = ix->table[i].slot;
            if (heartbeat) *heartbeat = ix->table[i].heartbeat;
            return (max_age < 0 || now - ix->table[i].heartbeat <= max_age) ? 1 : 0;
        }
        i = (i + 1) & (ix->table.size() - 1);
    }
    return 0;
}

}  // extern "C"
