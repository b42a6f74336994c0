// oracle/ref_shim/boost -- TEST INFRASTRUCTURE, not product code and not Boost.
// Boost.Thread / Boost.Bind / Boost.Function names used by the reference's hot-path sources
// (util/IndexThreadReduce.h, DataStructures/Frame.h, FrameMemory.h, Tracking/TrackingReference.h, IOWrapper/Timestamp.h)
// mapped onto the C++17 standard library, so that those files compile unmodified into oracle/_ref/.
// Boost is on that path for locking and worker threads only, never for arithmetic.
#ifndef LSD_REF_SHIM_BOOST
#define LSD_REF_SHIM_BOOST
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <shared_mutex>
#include <thread>

namespace boost {
typedef std::mutex mutex;
typedef std::recursive_mutex recursive_mutex;
typedef std::condition_variable condition_variable;
typedef std::thread thread;
template<typename M> using unique_lock = std::unique_lock<M>;
template<typename M> using shared_lock = std::shared_lock<M>;
template<typename M> using lock_guard = std::lock_guard<M>;
template<typename Sig> using function = std::function<Sig>;
using std::bind;
using std::ref;
using std::cref;

namespace posix_time {
typedef std::chrono::milliseconds milliseconds;
typedef std::chrono::microseconds microseconds;
typedef std::chrono::seconds seconds;
}

class shared_mutex {
public:
    void lock() { m_.lock(); }
    bool try_lock() { return m_.try_lock(); }
    void unlock() { m_.unlock(); }
    void lock_shared() { m_.lock_shared(); }
    bool try_lock_shared() { return m_.try_lock_shared(); }
    void unlock_shared() { m_.unlock_shared(); }
    template<typename Rep, typename Period> bool timed_lock(const std::chrono::duration<Rep, Period>& d) { return m_.try_lock_for(d); }
private:
    std::shared_timed_mutex m_;
};

namespace this_thread {
template<typename Rep, typename Period> inline void sleep(const std::chrono::duration<Rep, Period>& d) { std::this_thread::sleep_for(d); }
}
}  // namespace boost

// boost/bind.hpp puts _1 ... _9 into the global namespace
using namespace std::placeholders;
#endif
