// oracle/ref/blaze/Math.h -- TEST INFRASTRUCTURE.  A small stand-in for the parts of blaze-lib (3rdparty/blaze, an empty
// submodule in /root/reference; pinned commit 89ee9476...) that the reference's MCTS sources use, so that node.cpp,
// nodedata.cpp, searchthread.cpp, mctsagent.cpp, evalinfo.cpp and util/blazeutil.h compile unchanged from where they lie
// (oracle/Makefile target `ref`).  Written from blaze's documented behaviour, not from its source:
//   * DynamicVector<T>: resize / reserve / extend / size / data / iterators / element access; converting copies
//   * element-wise + - * / between vectors and scalars with the usual C++ arithmetic conversions per element
//   * argmax / argmin = index of the FIRST extreme element; max / min / sum (sequential) / softmax / pow / sqrt
//   * subvector(v, first, n): a view
//   * blaze's expression RESTRUCTURING  (v * s) * w  ->  (v * w) * s   (DVecScalarMultExpr.h, "restructuring binary
//     arithmetic operators"): a scalar-vector product is kept symbolic until it is used, exactly so that
//     get_current_u_values (node.cpp:1061) evaluates element i as (P_i * (sqrt(N)/(n_i+1))) * cput like blaze does.
// Eager evaluation otherwise: results are DynamicVector of the common element type (blaze's expression templates
// evaluate element by element with the same types, so values agree; SIMD reductions of blaze::sum are NOT modelled --
// the sum here is sequential).
#pragma once
// (blaze/Math.h drags most of the standard library in; the reference's sources rely on that)
#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <limits>
#include <type_traits>
#include <vector>

namespace blaze {

template <typename D>
struct Vec {  // CRTP base of everything with size() and operator[]
    const D& self() const { return static_cast<const D&>(*this); }
};

template <typename T>
class DynamicVector : public Vec<DynamicVector<T>> {
   public:
    using ElementType = T;
    using Iterator = T*;
    using ConstIterator = const T*;
    DynamicVector() = default;
    explicit DynamicVector(size_t n) : v_(n) {}
    DynamicVector(size_t n, const T& init) : v_(n, init) {}
    DynamicVector(const DynamicVector&) = default;
    DynamicVector(DynamicVector&&) = default;
    template <typename D>
    DynamicVector(const Vec<D>& o) {  // converting copy / evaluation of a view or symbolic product
        assign(o.self());
    }
    DynamicVector& operator=(const DynamicVector&) = default;
    DynamicVector& operator=(DynamicVector&&) = default;
    template <typename D>
    DynamicVector& operator=(const Vec<D>& o) {
        assign(o.self());
        return *this;
    }
    DynamicVector& operator=(const T& value) {
        std::fill(v_.begin(), v_.end(), value);
        return *this;
    }
    size_t size() const { return v_.size(); }
    size_t capacity() const { return v_.capacity(); }
    void resize(size_t n, bool preserve = true) {
        (void)preserve;
        v_.resize(n);
    }
    void reserve(size_t n) { v_.reserve(n); }
    void extend(size_t n, bool preserve = true) {
        (void)preserve;
        v_.resize(v_.size() + n);
    }
    void clear() { v_.clear(); }
    T& operator[](size_t i) {
        assert(i < v_.size());
        return v_[i];
    }
    const T& operator[](size_t i) const {
        assert(i < v_.size());
        return v_[i];
    }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }
    T* begin() { return v_.data(); }
    T* end() { return v_.data() + v_.size(); }
    const T* begin() const { return v_.data(); }
    const T* end() const { return v_.data() + v_.size(); }
    const T* cbegin() const { return begin(); }
    const T* cend() const { return end(); }
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
    DynamicVector& operator*=(S s) {
        for (auto& x : v_) x = static_cast<T>(x * s);
        return *this;
    }
    template <typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
    DynamicVector& operator/=(S s) {
        for (auto& x : v_) x = static_cast<T>(x / s);
        return *this;
    }
    template <typename D>
    DynamicVector& operator+=(const Vec<D>& o) {
        for (size_t i = 0; i < v_.size(); ++i) v_[i] = static_cast<T>(v_[i] + o.self()[i]);
        return *this;
    }
    template <typename D>
    DynamicVector& operator-=(const Vec<D>& o) {
        for (size_t i = 0; i < v_.size(); ++i) v_[i] = static_cast<T>(v_[i] - o.self()[i]);
        return *this;
    }
    template <typename D>
    DynamicVector& operator*=(const Vec<D>& o) {
        for (size_t i = 0; i < v_.size(); ++i) v_[i] = static_cast<T>(v_[i] * o.self()[i]);
        return *this;
    }

   private:
    template <typename D>
    void assign(const D& o) {
        v_.resize(o.size());
        for (size_t i = 0; i < v_.size(); ++i) v_[i] = static_cast<T>(o[i]);
    }
    std::vector<T> v_;
};

template <typename T, size_t N, bool TF = false>
using StaticVector = DynamicVector<T>;  // only named, never used with its size semantics, by the compiled sources
template <typename T, size_t N, bool TF = false>
using HybridVector = DynamicVector<T>;

// view on v[first, first+n)
template <typename V>
class Subvector : public Vec<Subvector<V>> {
   public:
    using ElementType = typename V::ElementType;
    Subvector(V& v, size_t first, size_t n) : v_(&v), first_(first), n_(n) { assert(first + n <= v.size()); }
    size_t size() const { return n_; }
    decltype(auto) operator[](size_t i) const { return (*v_)[first_ + i]; }
    decltype(auto) operator[](size_t i) { return (*v_)[first_ + i]; }

   private:
    V* v_;
    size_t first_, n_;
};
template <typename T>
Subvector<DynamicVector<T>> subvector(DynamicVector<T>& v, size_t first, size_t n) {
    return Subvector<DynamicVector<T>>(v, first, n);
}
template <typename T>
Subvector<const DynamicVector<T>> subvector(const DynamicVector<T>& v, size_t first, size_t n) {
    return Subvector<const DynamicVector<T>>(v, first, n);
}

template <typename D>
using Elem = std::decay_t<decltype(std::declval<const D&>()[0])>;

// symbolic (vector * scalar): element i = v[i] * s, with s converted to the common type like blaze's DVecScalarMultExpr
template <typename T, typename S>
class ScalarMult : public Vec<ScalarMult<T, S>> {
   public:
    using ElementType = std::common_type_t<T, S>;
    ScalarMult(DynamicVector<T> v, S s) : v_(std::move(v)), s_(s) {}
    size_t size() const { return v_.size(); }
    ElementType operator[](size_t i) const { return static_cast<ElementType>(v_[i]) * static_cast<ElementType>(s_); }
    const DynamicVector<T>& vector() const { return v_; }
    S scalar() const { return s_; }

   private:
    DynamicVector<T> v_;
    S s_;
};

#define BLAZE_SHIM_BINARY(op)                                                                                          \
    template <typename A, typename B>                                                                                  \
    DynamicVector<std::common_type_t<Elem<A>, Elem<B>>> operator op(const Vec<A>& a, const Vec<B>& b) {                \
        using R = std::common_type_t<Elem<A>, Elem<B>>;                                                                \
        assert(a.self().size() == b.self().size());                                                                    \
        DynamicVector<R> r(a.self().size());                                                                           \
        for (size_t i = 0; i < r.size(); ++i) r[i] = static_cast<R>(a.self()[i]) op static_cast<R>(b.self()[i]);       \
        return r;                                                                                                      \
    }
BLAZE_SHIM_BINARY(+)
BLAZE_SHIM_BINARY(-)
BLAZE_SHIM_BINARY(/)
#undef BLAZE_SHIM_BINARY

template <typename A, typename B>
DynamicVector<std::common_type_t<Elem<A>, Elem<B>>> mult(const A& a, const B& b) {
    using R = std::common_type_t<Elem<A>, Elem<B>>;
    assert(a.size() == b.size());
    DynamicVector<R> r(a.size());
    for (size_t i = 0; i < r.size(); ++i) r[i] = static_cast<R>(a[i]) * static_cast<R>(b[i]);
    return r;
}
// plain element-wise product
template <typename A, typename B>
auto operator*(const Vec<A>& a, const Vec<B>& b) {
    return mult(a.self(), b.self());
}
// restructuring: (v * s) * w -> (v * w) * s   and   w * (v * s) -> (w * v) * s
template <typename T, typename S, typename B>
auto operator*(const ScalarMult<T, S>& a, const Vec<B>& b) {
    auto vw = mult(a.vector(), b.self());
    return ScalarMult<Elem<decltype(vw)>, S>(std::move(vw), a.scalar());
}
template <typename A, typename T, typename S>
auto operator*(const Vec<A>& a, const ScalarMult<T, S>& b) {
    auto vw = mult(a.self(), b.vector());
    return ScalarMult<Elem<decltype(vw)>, S>(std::move(vw), b.scalar());
}

// vector (op) scalar and scalar (op) vector
template <typename A, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
ScalarMult<Elem<A>, S> operator*(const Vec<A>& a, S s) {
    return ScalarMult<Elem<A>, S>(DynamicVector<Elem<A>>(a), s);
}
template <typename A, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>
ScalarMult<Elem<A>, S> operator*(S s, const Vec<A>& a) {
    return ScalarMult<Elem<A>, S>(DynamicVector<Elem<A>>(a), s);
}
#define BLAZE_SHIM_SCALAR(op)                                                                                          \
    template <typename A, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>                       \
    DynamicVector<std::common_type_t<Elem<A>, S>> operator op(const Vec<A>& a, S s) {                                  \
        using R = std::common_type_t<Elem<A>, S>;                                                                      \
        DynamicVector<R> r(a.self().size());                                                                           \
        for (size_t i = 0; i < r.size(); ++i) r[i] = static_cast<R>(a.self()[i]) op static_cast<R>(s);                 \
        return r;                                                                                                      \
    }                                                                                                                  \
    template <typename A, typename S, typename = std::enable_if_t<std::is_arithmetic<S>::value>>                       \
    DynamicVector<std::common_type_t<Elem<A>, S>> operator op(S s, const Vec<A>& a) {                                  \
        using R = std::common_type_t<Elem<A>, S>;                                                                      \
        DynamicVector<R> r(a.self().size());                                                                           \
        for (size_t i = 0; i < r.size(); ++i) r[i] = static_cast<R>(s) op static_cast<R>(a.self()[i]);                 \
        return r;                                                                                                      \
    }
BLAZE_SHIM_SCALAR(+)
BLAZE_SHIM_SCALAR(-)
BLAZE_SHIM_SCALAR(/)
#undef BLAZE_SHIM_SCALAR

template <typename A>
size_t argmax(const Vec<A>& a) {  // first maximum
    const A& v = a.self();
    if (v.size() < 2) return 0;
    size_t idx = 0;
    auto best = v[0];
    for (size_t i = 1; i < v.size(); ++i)
        if (best < v[i]) best = v[i], idx = i;
    return idx;
}
template <typename A>
size_t argmin(const Vec<A>& a) {
    const A& v = a.self();
    if (v.size() < 2) return 0;
    size_t idx = 0;
    auto best = v[0];
    for (size_t i = 1; i < v.size(); ++i)
        if (v[i] < best) best = v[i], idx = i;
    return idx;
}
template <typename A>
Elem<A> max(const Vec<A>& a) {
    const A& v = a.self();
    Elem<A> m = v[0];
    for (size_t i = 1; i < v.size(); ++i) m = v[i] > m ? v[i] : m;
    return m;
}
template <typename A>
Elem<A> min(const Vec<A>& a) {
    const A& v = a.self();
    Elem<A> m = v[0];
    for (size_t i = 1; i < v.size(); ++i) m = v[i] < m ? v[i] : m;
    return m;
}
template <typename A>
Elem<A> sum(const Vec<A>& a) {
    const A& v = a.self();
    Elem<A> s = Elem<A>();
    for (size_t i = 0; i < v.size(); ++i) s = static_cast<Elem<A>>(s + v[i]);
    return s;
}
template <typename A, typename S>
DynamicVector<Elem<A>> pow(const Vec<A>& a, S e) {  // element-wise std::pow (float ^ float -> powf)
    DynamicVector<Elem<A>> r(a.self().size());
    for (size_t i = 0; i < r.size(); ++i) r[i] = static_cast<Elem<A>>(std::pow(a.self()[i], static_cast<Elem<A>>(e)));
    return r;
}
template <typename A>
DynamicVector<Elem<A>> sqrt(const Vec<A>& a) {
    DynamicVector<Elem<A>> r(a.self().size());
    for (size_t i = 0; i < r.size(); ++i) r[i] = std::sqrt(a.self()[i]);
    return r;
}
template <typename A>
DynamicVector<Elem<A>> exp(const Vec<A>& a) {
    DynamicVector<Elem<A>> r(a.self().size());
    for (size_t i = 0; i < r.size(); ++i) r[i] = std::exp(a.self()[i]);
    return r;
}
template <typename A>
DynamicVector<Elem<A>> softmax(const Vec<A>& a) {  // blaze: tmp = exp(v - max(v)); tmp / sum(tmp)
    DynamicVector<Elem<A>> r(a.self().size());
    if (r.size() == 0) return r;
    const Elem<A> m = max(a);
    for (size_t i = 0; i < r.size(); ++i) r[i] = std::exp(a.self()[i] - m);
    const Elem<A> s = sum(r);
    for (size_t i = 0; i < r.size(); ++i) r[i] /= s;
    return r;
}

}  // namespace blaze
