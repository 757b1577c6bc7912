#include <common/lcm_log.hpp>
#include <cstring>

namespace {

const uint32_t kSync = 0xEDA1DA01u;

bool readExact(FILE* f, void* dst, size_t n) { return std::fread(dst, 1, n, f) == n; }

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint64_t be64(const uint8_t* p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
float beFloat(const uint8_t* p) { const uint32_t u = be32(p); float f; std::memcpy(&f, &u, 4); return f; }

// lcmgen.c: hash_update / hash_string_update
int64_t hashUpdate(int64_t v, char c)
{
    // lcmgen.c: v = ((v<<8) ^ (v>>55)) + c on an int64_t: the right shift of a negative value is arithmetic (gcc)
    const int64_t left = (int64_t)((uint64_t)v << 8);
    const int64_t right = v >> 55;
    return (int64_t)((uint64_t)(left ^ right) + (uint64_t)(int64_t)c);
}
int64_t hashStringUpdate(int64_t v, const char* s)
{
    v = hashUpdate(v, (char)std::strlen(s));
    for (; *s != 0; ++s) v = hashUpdate(v, *s);
    return v;
}

}  // namespace

bool LcmLogReader::open(const std::string& path)
{
    close();
    file_ = std::fopen(path.c_str(), "rb");
    resyncs_ = 0;
    return file_ != nullptr;
}

void LcmLogReader::close(void)
{
    if (file_) std::fclose(file_);
    file_ = nullptr;
}

bool LcmLogReader::next(LcmLogEvent& ev)
{
    if (!file_) return false;
    for (;;) {
        // find the sync word (byte by byte after damage, like lcm_eventlog_read_next_event)
        uint32_t window = 0;
        int got = 0;
        bool skipped = false;
        for (;;) {
            const int c = std::fgetc(file_);
            if (c == EOF) return false;
            window = (window << 8) | (uint32_t)c;
            if (++got >= 4) {
                if (window == kSync) break;
                skipped = true;
            }
        }
        if (skipped) ++resyncs_;
        uint8_t head[24];
        if (!readExact(file_, head, sizeof(head))) return false;
        ev.eventNumber = (int64_t)be64(head);
        ev.timestamp = (int64_t)be64(head + 8);
        const int32_t chanLen = (int32_t)be32(head + 16), dataLen = (int32_t)be32(head + 20);
        if (chanLen <= 0 || chanLen >= 1000 || dataLen < 0 || dataLen > (64 << 20)) { ++resyncs_; continue; }   // LCM's own sanity limits
        ev.channel.resize((size_t)chanLen);
        if (!readExact(file_, &ev.channel[0], (size_t)chanLen)) return false;
        ev.data.resize((size_t)dataLen);
        if (dataLen > 0 && !readExact(file_, ev.data.data(), (size_t)dataLen)) return false;
        return true;
    }
}

int64_t lcmFingerprint(const std::vector<LcmMember>& members)
{
    int64_t v = 0x12345678;
    for (size_t i = 0; i < members.size(); ++i) {
        const LcmMember& m = members[i];
        v = hashStringUpdate(v, m.name);
        v = hashStringUpdate(v, m.type);                  // primitive member types hash their name
        v = hashUpdate(v, (char)m.dims.size());
        for (size_t d = 0; d < m.dims.size(); ++d) {
            v = hashUpdate(v, (char)m.dimModes[d]);       // 0 = constant size, 1 = variable (sized by a field)
            v = hashStringUpdate(v, m.dims[d].c_str());
        }
    }
    // __lcm_hash_recursive: no nested types to add; rotate left by one
    const uint64_t u = (uint64_t)v;
    return (int64_t)((u << 1) + ((u >> 63) & 1));
}

int64_t lidarFingerprint(void)
{
    std::vector<LcmMember> m;
    m.push_back(LcmMember{"utime", "int64_t", {}, {}});
    m.push_back(LcmMember{"num_ranges", "int32_t", {}, {}});
    m.push_back(LcmMember{"ranges", "float", {"num_ranges"}, {1}});
    m.push_back(LcmMember{"thetas", "float", {"num_ranges"}, {1}});
    m.push_back(LcmMember{"times", "int64_t", {"num_ranges"}, {1}});
    m.push_back(LcmMember{"intensities", "float", {"num_ranges"}, {1}});
    return lcmFingerprint(m);
}

int64_t odometryFingerprint(void)
{
    std::vector<LcmMember> m;
    m.push_back(LcmMember{"utime", "int64_t", {}, {}});
    m.push_back(LcmMember{"x", "float", {}, {}});
    m.push_back(LcmMember{"y", "float", {}, {}});
    m.push_back(LcmMember{"theta", "float", {}, {}});
    return lcmFingerprint(m);
}

bool decodeLidar(const std::vector<uint8_t>& d, lidar_t& scan, int64_t* fp)
{
    if (d.size() < 20) return false;
    if (fp) *fp = (int64_t)be64(d.data());
    scan.utime = (int64_t)be64(d.data() + 8);
    const int32_t n = (int32_t)be32(d.data() + 16);
    if (n < 0 || d.size() != 20 + (size_t)n * 20) return false;
    scan.num_ranges = n;
    scan.ranges.resize((size_t)n); scan.thetas.resize((size_t)n); scan.times.resize((size_t)n); scan.intensities.resize((size_t)n);
    const uint8_t* p = d.data() + 20;
    for (int32_t i = 0; i < n; ++i, p += 4) scan.ranges[(size_t)i] = beFloat(p);
    for (int32_t i = 0; i < n; ++i, p += 4) scan.thetas[(size_t)i] = beFloat(p);
    for (int32_t i = 0; i < n; ++i, p += 8) scan.times[(size_t)i] = (int64_t)be64(p);
    for (int32_t i = 0; i < n; ++i, p += 4) scan.intensities[(size_t)i] = beFloat(p);
    return true;
}

bool decodeOdometry(const std::vector<uint8_t>& d, pose_xyt_t& o, int64_t* fp)
{
    if (d.size() != 28) return false;
    if (fp) *fp = (int64_t)be64(d.data());
    o.utime = (int64_t)be64(d.data() + 8);
    o.x = beFloat(d.data() + 16); o.y = beFloat(d.data() + 20); o.theta = beFloat(d.data() + 24);
    return true;
}
