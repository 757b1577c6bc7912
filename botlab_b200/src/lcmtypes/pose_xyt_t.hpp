// Stand-in for the lcm-gen output of lcmtypes/pose_xyt_t.lcm (lcm-gen is not in this image).  Inside a botLab checkout
// the generated header takes this file's place: same class name, same public fields, same 24-byte layout.
#ifndef B200_LCMTYPES_POSE_XYT_T_HPP
#define B200_LCMTYPES_POSE_XYT_T_HPP
#include <cstdint>
class pose_xyt_t
{
public:
    int64_t utime = 0;
    float x = 0.0f;
    float y = 0.0f;
    float theta = 0.0f;
};
#endif
