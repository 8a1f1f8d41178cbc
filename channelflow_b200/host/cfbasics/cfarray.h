// cfarray<T>: the small 1-d container of the Channelflow API (reference cfbasics/cfarray.h), a thin shell over
// std::vector with the reference's member names.
#ifndef CFB200_CFARRAY_H
#define CFB200_CFARRAY_H
#include <algorithm>
#include <cassert>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

namespace chflow {

typedef double Real;
const int REAL_OUTPUT_DIGITS = 17;

template <class T>
class cfarray {
   public:
    using data_container_t = std::vector<T>;
    cfarray(int N = 0) : v_(N) {}
    cfarray(int N, const T& t) : v_(N, t) {}
    bool operator==(const cfarray& a) { return v_ == a.v_; }
    bool operator!=(const cfarray& a) { return !(v_ == a.v_); }
    void resize(int N) { v_.resize(N); }
    void fill(const T& t) { std::fill(v_.begin(), v_.end(), t); }
    typename data_container_t::reference operator[](int i) { assert(i >= 0 && (size_t)i < v_.size()); return v_[i]; }
    typename data_container_t::const_reference operator[](int i) const { assert(i >= 0 && (size_t)i < v_.size()); return v_[i]; }
    cfarray subvector(int offset, int N) const {
        cfarray s(N);
        std::copy(v_.begin() + offset, v_.begin() + offset + N, s.v_.begin());
        return s;
    }
    int N() const { return (int)v_.size(); }
    int length() const { return (int)v_.size(); }
    const T* pointer() const { return v_.data(); }
    T* pointer() { return v_.data(); }
    void save(const std::string& filebase) const {  // "% N" header, one element per line (the reference's .asc form)
        std::ofstream os((filebase + ".asc").c_str());
        os << std::scientific << std::setprecision(REAL_OUTPUT_DIGITS) << "% " << v_.size() << '\n';
        for (const auto& x : v_) os << x << '\n';
    }

   private:
    std::vector<T> v_;
};

template <class T>
std::ostream& operator<<(std::ostream& os, const cfarray<T>& a) {
    for (int i = 0; i < a.length(); ++i) os << a[i] << (a.length() < 10 ? ' ' : '\n');
    return os;
}

}  // namespace chflow
#endif
