// LCM event-log reader for the headless replay driver: `slam --localization-only` is fed by
// `lcm-logplayer` from a .log recorded with `lcm-logger` (reference README / data/*.log); this image has no LCM, so the
// log's framing and the two message types the SLAM loop consumes are decoded here.
//
// Event framing (lcm/eventlog.c, LCM 1.4.0 -- the version the reference's docker/Dockerfile:30 pins), all big-endian:
//   u32 sync = 0xEDA1DA01 | i64 event number | i64 timestamp (us) | i32 channel length | i32 data length | channel | data
// Message payload (lcm-gen C++ codecs): i64 fingerprint, then the fields in declaration order, big-endian; variable
// arrays are preceded by nothing (their length is an earlier field).  lidar_t: lcmtypes/lidar_t.lcm:1-14;
// odometry_t: lcmtypes/odometry_t.lcm:1-9 (same fields as pose_xyt_t).
#ifndef B200_COMMON_LCM_LOG_HPP
#define B200_COMMON_LCM_LOG_HPP

#include <lcmtypes/lidar_t.hpp>
#include <lcmtypes/pose_xyt_t.hpp>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

struct LcmLogEvent
{
    int64_t eventNumber;
    int64_t timestamp;
    std::string channel;
    std::vector<uint8_t> data;
};

class LcmLogReader
{
public:
    LcmLogReader(void) : file_(nullptr), resyncs_(0) {}
    ~LcmLogReader(void) { close(); }
    bool open(const std::string& path);
    void close(void);
    /// Next event, false at end of file.  A damaged region is skipped by scanning for the next sync word (like
    /// lcm_eventlog_read_next_event); resyncs() counts how often that happened.
    bool next(LcmLogEvent& event);
    int resyncs(void) const { return resyncs_; }

private:
    FILE* file_;
    int resyncs_;
};

/// The fingerprint lcm-gen assigns to a type made of primitive members only (lcmgen.c: lcm_struct_hash +
/// __lcm_hash_recursive's final rotation).  members: {name, type, dimension sizes (a field name for variable arrays)}.
struct LcmMember { const char* name; const char* type; std::vector<std::string> dims; std::vector<int> dimModes; };
int64_t lcmFingerprint(const std::vector<LcmMember>& members);
int64_t lidarFingerprint(void);
int64_t odometryFingerprint(void);

/// Decoders.  They check the payload's length against its own num_ranges and return false on any inconsistency;
/// fingerprintOut (may be null) receives the fingerprint found, which callers may compare with *Fingerprint().
bool decodeLidar(const std::vector<uint8_t>& data, lidar_t& scan, int64_t* fingerprintOut);
bool decodeOdometry(const std::vector<uint8_t>& data, pose_xyt_t& odometry, int64_t* fingerprintOut);

#endif  // B200_COMMON_LCM_LOG_HPP
