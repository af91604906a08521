// Minimal stand-in for boost::posix_time as used by utils/time_system.h:3,12-33 and
// actor/zero_actor.cpp:39-43. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <chrono>
#include <cstdint>
#include <ctime>

namespace boost {
namespace posix_time {

class time_duration {
public:
    explicit time_duration(int64_t us = 0) : us_(us) {}
    int64_t hours() const { return us_ / 3600000000LL; }
    int64_t minutes() const { return (us_ / 60000000LL) % 60; }
    int64_t seconds() const { return (us_ / 1000000LL) % 60; }
    int64_t total_milliseconds() const { return us_ / 1000; }
    int64_t total_microseconds() const { return us_; }

private:
    int64_t us_;
};

class ptime {
public:
    struct date_type {
        int y, m, d;
        int year() const { return y; }
        int month() const { return m; }
        int day() const { return d; }
    };
    explicit ptime(int64_t us_since_epoch = 0) : us_(us_since_epoch) {}
    date_type date() const
    {
        std::time_t t = static_cast<std::time_t>(us_ / 1000000LL);
        std::tm tmv;
        localtime_r(&t, &tmv);
        return {tmv.tm_year + 1900, tmv.tm_mon + 1, tmv.tm_mday};
    }
    time_duration time_of_day() const
    {
        std::time_t t = static_cast<std::time_t>(us_ / 1000000LL);
        std::tm tmv;
        localtime_r(&t, &tmv);
        int64_t us = (static_cast<int64_t>(tmv.tm_hour) * 3600 + tmv.tm_min * 60 + tmv.tm_sec) * 1000000LL + us_ % 1000000LL;
        return time_duration(us);
    }
    time_duration operator-(const ptime& rhs) const { return time_duration(us_ - rhs.us_); }

private:
    int64_t us_;
};

struct microsec_clock {
    static ptime local_time()
    {
        auto now = std::chrono::system_clock::now().time_since_epoch();
        return ptime(std::chrono::duration_cast<std::chrono::microseconds>(now).count());
    }
};

} // namespace posix_time
} // namespace boost
