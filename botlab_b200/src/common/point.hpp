// Minimal 2-D point with the member names the reference's common/point.hpp exposes (x, y, converting constructor).
#ifndef B200_COMMON_POINT_HPP
#define B200_COMMON_POINT_HPP
template <typename T>
class Point
{
public:
    T x;
    T y;
    Point() : x(0), y(0) {}
    Point(T px, T py) : x(px), y(py) {}
    template <typename U>
    Point(const Point<U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
};
#endif
