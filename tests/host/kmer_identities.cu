// Host-side check of the word-parallel identities the CUDA kernels rely on, using the SAME __host__ __device__
// functions (genomix_b200/csrc/gx_internal.cuh): packing, reverse complement, canonical choice, edge-mask bits and the
// reconstruction of neighbour k-mers from (key, mask). Reads `k<TAB>read` lines from stdin, aggregates (key -> count, mask)
// exactly like extract_kernel + table_upsert would, and prints one line per node:
//     <key bytes hex> <count> <FF list> <FR list> <RF list> <RR list>      (lists: comma separated neighbour byte hex, ACGT order)
// tests/test_host_identities.py compares that with the oracle's graph. No GPU involved: this is NOT a product path.
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../genomix_b200/csrc/gx_internal.cuh"

using namespace gx;

template <int KW>
static void pack(const char* s, int k, u64 (&f)[KW]) {
    for (int i = 0; i < KW; ++i) f[i] = 0;
    for (int i = 0; i < k; ++i) {
        bool ok;
        const u64 c = code_of((unsigned char)s[i], ok);
        f[i >> 5] |= c << (2 * (i & 31));
    }
}

template <int KW>
static std::string key_hex(const u64 (&w)[KW], int k) {
    const int nb = (k + 3) / 4;
    std::string out;
    char buf[4];
    for (int j = 0; j < nb; ++j) {
        const int byte_idx = nb - 1 - j;
        snprintf(buf, sizeof buf, "%02x", (unsigned)((w[byte_idx >> 3] >> (8 * (byte_idx & 7))) & 0xff));
        out += buf;
    }
    return out;
}

template <int KW>
struct KeyLess {
    bool operator()(const std::vector<u64>& a, const std::vector<u64>& b) const { return a < b; }
};

template <int KW>
static void run(int k, const std::vector<std::string>& reads) {
    std::map<std::vector<u64>, std::pair<u64, u32>> table;  // key words -> (count, mask)
    for (const std::string& r : reads) {
        const int len = (int)r.size();
        const int npos = len - k + 1;
        if (npos < 2) continue;
        std::vector<char> rev(npos);
        std::vector<std::vector<u64>> keys(npos);
        for (int p = 0; p < npos; ++p) {
            u64 f[KW], rc[KW];
            pack<KW>(r.c_str() + p, k, f);
            revcomp_key<KW>(f, k, rc);
            rev[p] = !key_le<KW>(f, rc);
            keys[p].assign(rev[p] ? rc : f, (rev[p] ? rc : f) + KW);
        }
        for (int p = 0; p < npos; ++p) {
            u32 mask = 0;
            bool ok;
            if (p + 1 < npos) mask |= edge_bit_next(rev[p], rev[p + 1], code_of((unsigned char)r[p + k], ok));
            if (p > 0) mask |= edge_bit_prev(rev[p], rev[p - 1], code_of((unsigned char)r[p - 1], ok));
            auto& slot = table[keys[p]];
            slot.first += 1;
            slot.second |= mask;
        }
    }
    for (const auto& kv : table) {
        u64 key[KW];
        for (int i = 0; i < KW; ++i) key[i] = kv.first[i];
        printf("%s %llu", key_hex<KW>(key, k).c_str(), (unsigned long long)kv.second.first);
        for (int t = 0; t < 4; ++t) {
            std::string lst;
            for (u32 b = 0; b < 4; ++b) {
                if (!((kv.second.second >> (4 * t + b)) & 1u)) continue;
                u64 nk[KW];
                neighbour_key<KW>(key, k, t, b, nk);
                if (!lst.empty()) lst += ",";
                lst += key_hex<KW>(nk, k);
            }
            printf(" %s", lst.empty() ? "-" : lst.c_str());
        }
        printf("\n");
    }
}

int main() {
    int k = 0;
    std::vector<std::string> reads;
    char line[1 << 16];
    while (fgets(line, sizeof line, stdin)) {
        char* tab = strchr(line, '\t');
        if (!tab) continue;
        *tab = 0;
        k = atoi(line);
        std::string r(tab + 1);
        while (!r.empty() && (r.back() == '\n' || r.back() == '\r')) r.pop_back();
        reads.push_back(r);
    }
    switch ((k + 31) / 32) {
        case 1: run<1>(k, reads); break;
        case 2: run<2>(k, reads); break;
        case 3: run<3>(k, reads); break;
        case 4: run<4>(k, reads); break;
        default: return 2;
    }
    return 0;
}
