// Minimal stand-in for the parts of Boost.Thread the reference's actor path touches
// (utils/paralleler.h:3,44-45,75,82; actor/actor_group.cpp:161). TEST INFRASTRUCTURE ONLY:
// lets oracle/Makefile compile the unmodified reference sources without a Boost install.
#pragma once
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace boost {

class barrier {
public:
    explicit barrier(unsigned count) : threshold_(count), count_(count), generation_(0) {}
    bool wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        unsigned gen = generation_;
        if (--count_ == 0) {
            ++generation_;
            count_ = threshold_;
            cv_.notify_all();
            return true;
        }
        cv_.wait(lk, [&] { return gen != generation_; });
        return false;
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    unsigned threshold_, count_, generation_;
};

using std::bind;
typedef std::mutex mutex;
template <class M>
using lock_guard = std::lock_guard<M>;

class thread_group {
public:
    template <class F>
    std::thread* create_thread(F f)
    {
        threads_.emplace_back(new std::thread(f));
        return threads_.back().get();
    }
    void interrupt_all() {}
    void join_all()
    {
        // the reference's slave threads never return (isDone() == false); detach instead of join
        for (auto& t : threads_) {
            if (t->joinable()) { t->detach(); }
        }
    }

private:
    std::vector<std::unique_ptr<std::thread>> threads_;
};

} // namespace boost
