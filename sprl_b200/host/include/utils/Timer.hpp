// utils/Timer.hpp -- the stopwatch the reference's mains use (cpp/src/utils/Timer.hpp: reset(), elapsed() in seconds; global
// namespace).  Time.cpp reaches it through evaluate/play.hpp, as in the reference.
#ifndef SPRL_B200_COMPAT_UTILS_TIMER_HPP
#define SPRL_B200_COMPAT_UTILS_TIMER_HPP

#include <chrono>

class Timer {
public:
    Timer() { reset(); }
    void reset() { m_start = std::chrono::steady_clock::now(); }
    double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - m_start).count(); }
private:
    std::chrono::steady_clock::time_point m_start;
};

#endif
