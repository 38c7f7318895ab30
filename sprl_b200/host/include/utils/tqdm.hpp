// utils/tqdm.hpp -- include-path compatibility with the reference's cpp/src/utils/tqdm.hpp (a third-party progress bar,
// out of scope): just enough for `for (int t : tq::trange(n))` and `pbar << "text"` in the reference's mains.
#ifndef SPRL_B200_COMPAT_UTILS_TQDM_HPP
#define SPRL_B200_COMPAT_UTILS_TQDM_HPP
#include <iostream>
#include <sstream>
#include <string>

namespace tq {

class plain_range {
public:
    explicit plain_range(long long n) : m_n(n) {}
    struct iterator {
        long long i; plain_range* owner;
        long long operator*() const { return i; }
        iterator& operator++() { owner->flush(); ++i; return *this; }
        bool operator!=(const iterator& o) const { return i != o.i; }
    };
    iterator begin() { return { 0, this }; }
    iterator end() { return { m_n, this }; }
    template <typename T> plain_range& operator<<(const T& t) { m_suffix << t; return *this; }
    void flush() {
        if (!m_suffix.str().empty()) std::cout << m_suffix.str() << std::endl;
        m_suffix.str(std::string());
    }
private:
    long long m_n;
    std::ostringstream m_suffix;
};

template <typename IntType> plain_range trange(IntType last) { return plain_range((long long)last); }

}  // namespace tq
#endif
