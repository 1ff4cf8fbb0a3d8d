// egpu_restore.cc — host side of placement-state restore (include/egpu_restore.h): reads the
// reference's stored formats and hands flat arrays to egpu_table_restore_flat (egpu_devhash.cu),
// where the identity check and the sums run on the GPU.
//
//   record key   "namespace/name"                      pkg/types/pod.go:39-44,51-53
//   record value json.Marshal(map[string]*types.Device) pkg/types/pod.go:45-47,55-58
//                Device = {Hash string; List []string; ResourceName v1.ResourceName}
//                                                       pkg/types/device.go:11-15
//   symlink      /host/dev/elastic-gpu-<Hash>-<i> -> /dev/nvidia<N>
//                                                       pkg/operator/gpushare.go:10-14,31-55
//
// The JSON reader accepts what encoding/json accepts for this shape: any whitespace, any key
// order, unknown keys skipped, key names matched case-insensitively, null for the map, an entry
// or the list, string escapes including \uXXXX surrogate pairs.  Anything else is
// EGPU_ERR_PARSE for the whole call (NewPIFromRaw's error aborts Storage.ForEach the same way).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <atomic>
#include <thread>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "../../include/egpu_restore.h"

// egpu_alloc.cu (internal, not in the public headers): records a message for egpu_last_error
void egpu_note_error(egpu_ctx* ctx, const char* msg);

namespace {

// One container's stored Device.  The IDs of List are kept the way the device wants them - the
// characters back to back plus the end offset of each - so that flattening a record is two bulk
// appends, not one small string per ID (a gpu-memory container holds one ID per MiB).
struct Entry {
    std::string container, hash, resource;
    std::string ids;             // List[0] List[1] ... without separators
    std::vector<uint32_t> ends;  // end offset of every ID in `ids`
    bool ids_wellformed = true;  // every ID is 1..16 characters of '-' and digits
    bool is_null = false;
    void clear_list() {
        ids.clear();
        ends.clear();
        ids_wellformed = true;
    }
    void add_id(const char* p, size_t n) {
        bool ok = n >= 1 && n <= 16 && ids.size() < 0xF0000000u;  // (the offsets are 32-bit per entry)
        for (size_t i = 0; i < n; ++i) ok = ok && (p[i] == '-' || (p[i] >= '0' && p[i] <= '9'));
        ids_wellformed = ids_wellformed && ok;
        if (n > 64) n = 64;  // malformed anyway (the call fails below); do not hoard its bytes
        ids.append(p, n);
        ends.push_back(static_cast<uint32_t>(ids.size()));
    }
};

class Json {
  public:
    Json(const char* p, int64_t n) : p_(p), end_(p + n) {}

    // {"container": null | {Hash, List, ResourceName}, ...} or null
    bool parse_record(std::vector<Entry>& out) {
        ws();
        if (lit("null")) return tail();
        if (!eat('{')) return false;
        std::map<std::string, Entry> by_name;  // a repeated key keeps the last value, as encoding/json does
        std::vector<std::string> order;
        ws();
        if (!eat('}')) {
            for (;;) {
                Entry e;
                ws();
                if (!str(e.container)) return false;
                ws();
                if (!eat(':')) return false;
                ws();
                if (lit("null")) e.is_null = true;
                else if (!device(e)) return false;
                if (!by_name.count(e.container)) order.push_back(e.container);
                by_name[e.container] = std::move(e);
                ws();
                if (eat(',')) continue;
                if (eat('}')) break;
                return false;
            }
        }
        for (const std::string& k : order) out.push_back(std::move(by_name[k]));
        return tail();
    }

  private:
    const char* p_;
    const char* end_;

    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
    }
    bool eat(char c) {
        if (p_ < end_ && *p_ == c) {
            ++p_;
            return true;
        }
        return false;
    }
    bool lit(const char* w) {
        const size_t l = std::strlen(w);
        if (static_cast<size_t>(end_ - p_) >= l && std::memcmp(p_, w, l) == 0) {
            p_ += l;
            return true;
        }
        return false;
    }
    bool tail() {
        ws();
        return p_ == end_;
    }
    static bool ieq(const std::string& a, const char* b) {
        const size_t l = std::strlen(b);
        if (a.size() != l) return false;
        for (size_t i = 0; i < l; ++i) {
            char x = a[i], y = b[i];
            if (x >= 'A' && x <= 'Z') x = static_cast<char>(x - 'A' + 'a');
            if (y >= 'A' && y <= 'Z') y = static_cast<char>(y - 'A' + 'a');
            if (x != y) return false;
        }
        return true;
    }
    bool hex4(unsigned& v) {
        if (end_ - p_ < 4) return false;
        v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = *p_++;
            const int d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1;
            if (d < 0) return false;
            v = v * 16 + static_cast<unsigned>(d);
        }
        return true;
    }
    static void utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o += static_cast<char>(cp);
        else if (cp < 0x800) o += static_cast<char>(0xC0 | (cp >> 6)), o += static_cast<char>(0x80 | (cp & 63));
        else if (cp < 0x10000)
            o += static_cast<char>(0xE0 | (cp >> 12)), o += static_cast<char>(0x80 | ((cp >> 6) & 63)), o += static_cast<char>(0x80 | (cp & 63));
        else
            o += static_cast<char>(0xF0 | (cp >> 18)), o += static_cast<char>(0x80 | ((cp >> 12) & 63)),
                o += static_cast<char>(0x80 | ((cp >> 6) & 63)), o += static_cast<char>(0x80 | (cp & 63));
    }
    bool str(std::string& o) {
        o.clear();
        if (!eat('"')) return false;
        {   // fast path (every device ID, hash and resource name): no escape before the closing quote
            const char* q = static_cast<const char*>(std::memchr(p_, '"', static_cast<size_t>(end_ - p_)));
            if (!q) return false;
            bool simple = true;
            for (const char* c = p_; c < q; ++c)
                if (*c == '\\' || static_cast<unsigned char>(*c) < 0x20) {
                    simple = false;
                    break;
                }
            if (simple) {
                o.assign(p_, q);
                p_ = q + 1;
                return true;
            }
        }
        while (p_ < end_) {
            const unsigned char c = static_cast<unsigned char>(*p_++);
            if (c == '"') return true;
            if (c < 0x20) return false;
            if (c != '\\') {
                o += static_cast<char>(c);
                continue;
            }
            if (p_ >= end_) return false;
            const char e = *p_++;
            switch (e) {
                case '"': o += '"'; break;
                case '\\': o += '\\'; break;
                case '/': o += '/'; break;
                case 'b': o += '\b'; break;
                case 'f': o += '\f'; break;
                case 'n': o += '\n'; break;
                case 'r': o += '\r'; break;
                case 't': o += '\t'; break;
                case 'u': {
                    unsigned cp;
                    if (!hex4(cp)) return false;
                    if (cp >= 0xD800 && cp < 0xDC00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        const char* save = p_;
                        p_ += 2;
                        unsigned lo;
                        if (!hex4(lo)) return false;
                        if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        else p_ = save, cp = 0xFFFD;
                    } else if (cp >= 0xD800 && cp < 0xE000) {
                        cp = 0xFFFD;  // lone surrogate: U+FFFD, as encoding/json
                    }
                    utf8(o, cp);
                    break;
                }
                default: return false;
            }
        }
        return false;
    }
    bool number() {
        const char* s = p_;
        eat('-');
        while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) ++p_;
        return p_ > s;
    }
    // any value, discarded (unknown keys)
    bool skip(int depth = 0) {
        if (depth > 64) return false;
        ws();
        if (p_ >= end_) return false;
        std::string tmp;
        switch (*p_) {
            case '"': return str(tmp);
            case '{':
                ++p_;
                ws();
                if (eat('}')) return true;
                for (;;) {
                    ws();
                    if (!str(tmp)) return false;
                    ws();
                    if (!eat(':') || !skip(depth + 1)) return false;
                    ws();
                    if (eat(',')) continue;
                    return eat('}');
                }
            case '[':
                ++p_;
                ws();
                if (eat(']')) return true;
                for (;;) {
                    if (!skip(depth + 1)) return false;
                    ws();
                    if (eat(',')) continue;
                    return eat(']');
                }
            case 't': return lit("true");
            case 'f': return lit("false");
            case 'n': return lit("null");
            default: return number();
        }
    }
    bool string_or_null(std::string& o) {
        if (lit("null")) {
            o.clear();
            return true;
        }
        return str(o);
    }
    bool device(Entry& e) {
        if (!eat('{')) return false;
        ws();
        if (eat('}')) return true;
        for (;;) {
            std::string k;
            ws();
            if (!str(k)) return false;
            ws();
            if (!eat(':')) return false;
            ws();
            if (ieq(k, "Hash")) {
                if (!string_or_null(e.hash)) return false;
            } else if (ieq(k, "ResourceName")) {
                if (!string_or_null(e.resource)) return false;
            } else if (ieq(k, "List")) {
                e.clear_list();
                if (!lit("null")) {
                    if (!eat('[')) return false;
                    ws();
                    if (!eat(']')) {
                        std::string id;
                        // one ID per MiB: size the buffers once (the rest of the value bounds both)
                        e.ids.reserve(static_cast<size_t>(end_ - p_));
                        e.ends.reserve(static_cast<size_t>(end_ - p_) / 4 + 1);
                        for (;;) {
                            ws();
                            // fast path: "<up to 16 of '-' and digits>" straight into the entry's buffers;
                            // with 20 bytes left the scan needs no bounds checks
                            bool taken = false;
                            if (end_ - p_ >= 20 && *p_ == '"' && e.ids.size() < 0xF0000000u) {
                                const char* b = p_ + 1;
                                const char* c = b;
                                while (c - b < 17 && (*c == '-' || (*c >= '0' && *c <= '9'))) ++c;
                                if (*c == '"' && c > b && c - b <= 16) {
                                    e.ids.append(b, static_cast<size_t>(c - b));
                                    e.ends.push_back(static_cast<uint32_t>(e.ids.size()));
                                    p_ = c + 1;
                                    taken = true;
                                }
                            }
                            if (!taken) {  // anything else: the general string reader, then the same checks
                                if (!str(id)) return false;
                                e.add_id(id.data(), id.size());
                            }
                            ws();
                            if (eat(',')) continue;
                            if (eat(']')) break;
                            return false;
                        }
                    }
                }
            } else if (!skip()) {
                return false;
            }
            ws();
            if (eat(',')) continue;
            return eat('}');
        }
    }
};

int fail(egpu_ctx* ctx, int code, const char* what, int64_t record) {
    char msg[200];
    std::snprintf(msg, sizeof msg, "egpu_table_restore: %s (record %lld)", what, static_cast<long long>(record));
    egpu_note_error(ctx, msg);
    return code;
}

// "elastic-gpu-<hash>-<i>" or "<hash>-<i>" -> (hash, i); false for anything else
bool split_link(const char* name, std::string& hash, int64_t& ordinal) {
    static const char kPrefix[] = "elastic-gpu-";
    std::string s(name);
    if (s.compare(0, sizeof kPrefix - 1, kPrefix) == 0) s.erase(0, sizeof kPrefix - 1);
    const size_t dash = s.rfind('-');
    if (dash == std::string::npos || dash == 0 || dash + 1 >= s.size() || s.size() - dash - 1 > 9) return false;
    int64_t v = 0;
    for (size_t i = dash + 1; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
    }
    hash = s.substr(0, dash);
    ordinal = v;
    return true;
}

// "/dev/nvidia<N>" (pkg/operator/gpushare.go:10) -> N; -1 for anything else
int32_t target_gpu(const char* target) {
    static const char kDev[] = "/dev/nvidia";
    if (std::strncmp(target, kDev, sizeof kDev - 1) != 0) return -1;
    const char* p = target + sizeof kDev - 1;
    if (!*p || std::strlen(p) > 6) return -1;
    int32_t v = 0;
    for (; *p; ++p) {
        if (*p < '0' || *p > '9') return -1;
        v = v * 10 + (*p - '0');
    }
    return v;
}

}  // namespace

extern "C" int egpu_table_restore(egpu_ctx* ctx, const char* const* keys, const int64_t* key_lens,
                                  const char* const* vals, const int64_t* val_lens, int64_t n_records,
                                  const char* const* link_names, const char* const* link_targets, int64_t n_links,
                                  const int32_t* cap_core, const int32_t* cap_mem, int32_t D, int flags,
                                  int32_t* out_table, int64_t* out_counts, int32_t* out_record_status) {
    if (!ctx || n_records < 0 || n_links < 0) return EGPU_ERR_INVALID;
    if (n_records > 0 && (!keys || !key_lens || !vals || !val_lens)) return EGPU_ERR_INVALID;
    if (n_links > 0 && (!link_names || !link_targets)) return EGPU_ERR_INVALID;

    // symlinks: hash -> GPU per ordinal
    std::map<std::string, std::vector<int32_t>> links;
    for (int64_t i = 0; i < n_links; ++i) {
        if (!link_names[i] || !link_targets[i]) return EGPU_ERR_INVALID;
        std::string hash;
        int64_t ord;
        if (std::strncmp(link_names[i], "elastic-gpuctl-", 15) == 0) continue;
        if (!split_link(link_names[i], hash, ord) || ord > 4096) continue;
        const int32_t gpu = target_gpu(link_targets[i]);
        if (gpu < 0) continue;
        std::vector<int32_t>& v = links[hash];
        if (static_cast<int64_t>(v.size()) <= ord) v.resize(static_cast<size_t>(ord) + 1, -1);
        v[static_cast<size_t>(ord)] = gpu;
    }

    // Phase 1, parallel over records: key check + JSON parse (the reader is most of the host time of a
    // node-scale restore: ~10 MB of record values).  Each record parses into its own slot; the merge
    // below walks the slots in record order, so results and the first reported error are those of a
    // sequential pass.
    for (int64_t r = 0; r < n_records; ++r)
        if (!keys[r] || key_lens[r] < 0 || val_lens[r] < 0 || (val_lens[r] > 0 && !vals[r])) return EGPU_ERR_INVALID;
    std::vector<std::vector<Entry>> parsed(static_cast<size_t>(n_records));
    std::vector<signed char> perr(static_cast<size_t>(n_records), 0);  // 1 = bad key, 2 = bad value
    {
        int64_t bytes = 0;
        for (int64_t r = 0; r < n_records; ++r) bytes += val_lens[r];
        int nthr = 1;
        if (const char* e = std::getenv("EGPU_RESTORE_THREADS")) nthr = std::atoi(e);
        else nthr = static_cast<int>(std::min<int64_t>(std::min<int64_t>(8, std::thread::hardware_concurrency()), bytes >> 18));  // >= 256 KB per thread
        if (nthr > n_records) nthr = static_cast<int>(n_records);
        if (nthr < 1) nthr = 1;
        std::atomic<int64_t> next{0};
        auto work = [&]() {
            for (;;) {
                const int64_t r = next.fetch_add(1, std::memory_order_relaxed);
                if (r >= n_records) return;
                // strings.Split(key, "/") must give exactly two parts (pkg/types/pod.go:40-43)
                int64_t slashes = 0;
                for (int64_t i = 0; i < key_lens[r]; ++i) slashes += keys[r][i] == '/';
                if (slashes != 1) {
                    perr[static_cast<size_t>(r)] = 1;
                    continue;
                }
                Json js(vals[r], val_lens[r]);
                if (!js.parse_record(parsed[static_cast<size_t>(r)])) perr[static_cast<size_t>(r)] = 2;
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthr; ++t) pool.emplace_back(work);
        work();
        for (std::thread& t : pool) t.join();
    }

    std::vector<char> flat;
    std::vector<int64_t> id_off{0}, set_off{0}, link_off{0};
    std::vector<char> hash8;
    std::vector<int32_t> resource, link_gpu;
    std::vector<int64_t> set_record;
    {
        size_t id_bytes = 0, n_id = 0;
        for (const std::vector<Entry>& es : parsed)
            for (const Entry& e : es) {
                id_bytes += e.ids.size();
                n_id += e.ends.size();
            }
        flat.reserve(id_bytes);
        id_off.reserve(n_id + 1);
    }
    for (int64_t r = 0; r < n_records; ++r) {
        if (perr[static_cast<size_t>(r)] == 1) return fail(ctx, EGPU_ERR_PARSE, "error key format", r);
        if (perr[static_cast<size_t>(r)] == 2) return fail(ctx, EGPU_ERR_PARSE, "error val format", r);
        std::vector<Entry>& entries = parsed[static_cast<size_t>(r)];
        for (Entry& e : entries) {
            int32_t res = EGPU_RESOURCE_FOREIGN;
            if (e.resource == "elasticgpu.io/gpu-core") res = EGPU_RESOURCE_CORE;
            else if (e.resource == "elasticgpu.io/gpu-memory") res = EGPU_RESOURCE_MEM;
            if (e.is_null) res = EGPU_RESOURCE_CORE, e.clear_list();  // a nil *Device holds nothing: reported EMPTY
            if (res != EGPU_RESOURCE_FOREIGN) {
                if (!e.ids_wellformed) return fail(ctx, EGPU_ERR_PARSE, "device ID is not \"<gpu>-<unit>\"", r);
                const int64_t base = static_cast<int64_t>(flat.size());
                flat.insert(flat.end(), e.ids.begin(), e.ids.end());
                for (uint32_t end : e.ends) id_off.push_back(base + static_cast<int64_t>(end));
            }
            set_off.push_back(static_cast<int64_t>(id_off.size()) - 1);
            char h[8];
            for (int k = 0; k < 8; ++k) h[k] = e.hash.size() == 8 ? e.hash[static_cast<size_t>(k)] : '?';
            hash8.insert(hash8.end(), h, h + 8);
            resource.push_back(res);
            auto it = links.find(e.hash);
            if (it != links.end()) link_gpu.insert(link_gpu.end(), it->second.begin(), it->second.end());
            link_off.push_back(static_cast<int64_t>(link_gpu.size()));
            set_record.push_back(r);
        }
    }
    const int64_t n_sets = static_cast<int64_t>(resource.size());
    const int64_t n_ids = static_cast<int64_t>(id_off.size()) - 1;
    std::vector<int32_t> status(static_cast<size_t>(n_sets), 0);
    const int rc = egpu_table_restore_flat(ctx, flat.data(), id_off.data(), n_ids, set_off.data(), n_sets, hash8.data(),
                                           resource.data(), link_off.data(), link_gpu.data(), cap_core, cap_mem, D, flags,
                                           out_table, status.data());
    if (rc != EGPU_OK) return rc;
    if (out_counts) {
        for (int k = 0; k < EGPU_REC_STATUS_COUNT; ++k) out_counts[k] = 0;
        for (int32_t st : status) out_counts[st] += 1;
    }
    if (out_record_status) {
        for (int64_t r = 0; r < n_records; ++r) out_record_status[r] = EGPU_REC_OK;
        for (int64_t q = 0; q < n_sets; ++q)
            if (status[static_cast<size_t>(q)] > out_record_status[set_record[static_cast<size_t>(q)]])
                out_record_status[set_record[static_cast<size_t>(q)]] = status[static_cast<size_t>(q)];
    }
    return EGPU_OK;
}
