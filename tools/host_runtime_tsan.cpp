// ThreadSanitizer driver for the C++ host runtime (csrc/host_runtime.cpp): hammers the thread-safe pieces from many
// threads — the expert index (put/get with growth), the Kademlia routing table (add/remove/closest) and the threaded
// gather.  Build + run (CPU only):
//   g++ -O1 -g -fsanitize=thread -std=c++17 -pthread tools/host_runtime_tsan.cpp learning-at-home_b200/csrc/host_runtime.cpp -o /tmp/host_tsan && /tmp/host_tsan
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {
int lah_host_gather(const void** srcs, const unsigned long long* nbytes, int n, void* dst, int threads);
void* lah_rt_create(const uint8_t* self_id, int k);
void lah_rt_destroy(void* h);
int lah_rt_add(void* h, const uint8_t* id, uint32_t ip, uint16_t port, double now, uint8_t* evict_id, uint32_t* evict_ip,
               uint16_t* evict_port);
int lah_rt_remove(void* h, const uint8_t* id);
int lah_rt_size(void* h);
int lah_rt_closest(void* h, const uint8_t* target, int n, uint8_t* out_ids, uint32_t* out_ips, uint16_t* out_ports);
unsigned long long lah_hash_bytes(const uint8_t* data, int n);
void* lah_index_create(int capacity_pow2);
void lah_index_destroy(void* h);
int lah_index_put(void* h, unsigned long long key, int owner, int slot, double heartbeat);
int lah_index_get(void* h, unsigned long long key, double now, double max_age, int* owner, int* slot, double* heartbeat);
}

static void make_id(uint8_t* id, uint32_t seed) {
    for (int i = 0; i < 20; ++i) {
        seed = seed * 1664525u + 1013904223u;
        id[i] = static_cast<uint8_t>(seed >> 24);
    }
}

int main() {
    const int T = 8, N = 4000;
    std::atomic<long> found{0}, errors{0};
    // ---- expert index: concurrent put / get while the table grows
    void* ix = lah_index_create(16);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (int i = 0; i < N; ++i) {
                    const unsigned long long key = 1 + static_cast<unsigned long long>(t) * N + i;
                    lah_index_put(ix, key, t, i, 1.0 + i);
                    int owner = -1, slot = -1;
                    double hb = 0;
                    if (lah_index_get(ix, key, 2.0 + i, -1.0, &owner, &slot, &hb)) {
                        ++found;
                        if (owner != t || slot != i) ++errors;
                    } else {
                        ++errors;
                    }
                    lah_index_get(ix, 1 + (key * 7919) % (static_cast<unsigned long long>(T) * N), 0.0, -1.0, &owner, &slot, &hb);
                }
            });
        for (auto& x : th) x.join();
    }
    lah_index_destroy(ix);
    // ---- routing table: concurrent add / remove / closest
    uint8_t self_id[20];
    make_id(self_id, 1);
    void* rt = lah_rt_create(self_id, 20);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                uint8_t id[20], ev_id[20], out_ids[20 * 8];
                uint32_t ev_ip, out_ips[8];
                uint16_t ev_port, out_ports[8];
                for (int i = 0; i < N; ++i) {
                    make_id(id, 1000 + t * N + i);
                    lah_rt_add(rt, id, 0x7f000001u, static_cast<uint16_t>(8000 + t), i * 0.001, ev_id, &ev_ip, &ev_port);
                    if (i % 3 == 0) lah_rt_closest(rt, id, 8, out_ids, out_ips, out_ports);
                    if (i % 5 == 0) lah_rt_remove(rt, id);
                    if (i % 97 == 0) lah_rt_size(rt);
                }
            });
        for (auto& x : th) x.join();
    }
    const int rt_size = lah_rt_size(rt);
    lah_rt_destroy(rt);
    // ---- threaded gather (each worker thread copies a disjoint byte range)
    const int parts = 64;
    std::vector<std::vector<char>> src(parts, std::vector<char>(1 << 16));
    std::vector<const void*> ptrs(parts);
    std::vector<unsigned long long> sizes(parts, 1 << 16);
    for (int i = 0; i < parts; ++i) {
        memset(src[i].data(), i, src[i].size());
        ptrs[i] = src[i].data();
    }
    std::vector<char> dst(static_cast<size_t>(parts) << 16);
    lah_host_gather(ptrs.data(), sizes.data(), parts, dst.data(), 8);
    for (int i = 0; i < parts; ++i)
        if (dst[(static_cast<size_t>(i) << 16) + 123] != static_cast<char>(i)) ++errors;
    printf("host_runtime_tsan: index hits %ld / %d, routing table size %d, logic errors %ld\n", found.load(), T * N, rt_size,
           errors.load());
    return errors.load() ? 1 : 0;
}
