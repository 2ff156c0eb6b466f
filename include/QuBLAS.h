// QuBLAS.h — B200-native drop-in for the fixed-point GEMM / tree-reduction hot path of
// pikapuma/QuBLAS.  Clean-room C++23 host header: it re-declares the part of the reference's
// template API that the path needs (same names, same tag grammar, same type rules) and lowers every
// call on tensors to the plain-C ABI of libqb200.so (include/qb200.h), where hand-written sm_100a
// kernels do the work.  Link with -lqb200.
//
//   reference API (file:line in the reference's include/QuBLAS.h)        here
//   -----------------------------------------------------------------    ---------------------------
//   intBits/fracBits/isSigned/QuMode/OfMode tags, defaults  :2346-2359    same names, same defaults
//   RND::* / TRN::* / SAT::* / WRP::TCPL ids                :1986-1999, :2209-2225   same ::value ids
//   Qu<tags...> scalar, converting ctor, toDouble, fill      :2368-2498    QuBLAS::Qu
//   Qcomplex<re,im>, Qu<re,im>                               :2501-2617    QuBLAS::Qcomplex
//   dim<...>, Qu<dim<...>, T> column-major tensor, m[i,j]    :2622-3049    same layout (raw words)
//   FullPrec, MulMerger / AddMerger result types             :3079-3140    detail::merge_*
//   Qmul / Qadd / Qsub on scalars                            :3142-3235, :3980-3999  constexpr, header-only
//   TFComplexMul<abT,cdT,baT,abcT,cdbT,badT,ABT,BCT>         :3453-3535    incl. its routing quirks
//   BasicComplexMul<acT,bdT,adT,bcT,acbdT,adbcT>, Qmul(c1,c2) :3376-3446    4 mult / 2 add (default complex product)
//   Qreduce<Args...>(tensor) / (scalars...)                  :4903-5018    tensor form -> GPU
//   Qgemul<QgemulAddArgs<>,QgemulMulArgs<>,QgemulTransposedA<>>(C,A,B)   readme.md:80-87 -> GPU
//   BitStream<l2r|r2l<n>...>, toString(), operator<<          :4545-4827, :2433, :2913   host-only wire format / printing
//   ANUS::Qpoly / Segment / Qapprox / Qapprox_s               :4829-4897    host-only, scalar
//
// Scalar Qmul/Qadd/Qsub/Qreduce(scalars...) are header-only conveniences so that existing user code
// keeps compiling; every tensor-level operation (Qreduce(tensor), Qgemul, tensor TFComplexMul) goes
// through the C ABI and fails with std::runtime_error when no CUDA device is present — there is no
// CPU fallback for the measured path.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <iomanip>
#include <iostream>
#include <new>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>
#include <vector>

#include "qb200.h"

inline namespace QuBLAS {

// ------------------------------------------------------------------------------------------------
// tags
// ------------------------------------------------------------------------------------------------
template <int V> struct intBits { static constexpr int value = V; };
template <int V> struct fracBits { static constexpr int value = V; };
template <bool V> struct isSigned { static constexpr bool value = V; };

struct RND {
    struct POS_INF { static constexpr int value = 0; };
    struct NEG_INF { static constexpr int value = 1; };
    struct ZERO { static constexpr int value = 2; };
    struct INF { static constexpr int value = 3; };
    struct CONV { static constexpr int value = 4; };
};
struct TRN {
    struct TCPL { static constexpr int value = 5; };
    struct SMGN { static constexpr int value = 6; };
};
struct SAT {
    struct TCPL { static constexpr int value = 0; };
    struct ZERO { static constexpr int value = 1; };
    struct SMGN { static constexpr int value = 2; };
};
struct WRP {
    struct TCPL { static constexpr int value = 3; };
};
template <class M> struct QuMode { using type = M; static constexpr int value = M::value; };
template <class M> struct OfMode { using type = M; static constexpr int value = M::value; };
struct FullPrec {};

template <class... Ts> struct TypeList { static constexpr std::size_t size = sizeof...(Ts); };

template <std::size_t... D> struct dim {
    static constexpr std::size_t dimSize = sizeof...(D);
    static constexpr std::size_t elemSize = (D * ... * 1);
    static constexpr std::array<std::size_t, sizeof...(D)> dimArray = { D... };
};

constexpr int defaultIntBits = 8;
constexpr int defaultFracBits = 8;
constexpr bool defaultIsSigned = true;
using defaultQuMode = TRN::TCPL;
using defaultOfMode = SAT::TCPL;

template <class... Args> class Qu_s;

namespace detail {

// ---- integer quantizer core (the same arithmetic the kernels run) ----
struct Spec { int i, f; bool s; int q, o; };
constexpr bool operator==(const Spec &a, const Spec &b) { return a.i == b.i && a.f == b.f && a.s == b.s && a.q == b.q && a.o == b.o; }

typedef __int128 i128;

constexpr i128 shl(i128 v, int s) { return v * ((i128)1 << s); }

constexpr i128 round_shift(i128 v, int s, int mode)
{
    if (s <= 0) return shl(v, -s);
    if (s > 120) return (mode == 5 && v < 0) ? -1 : 0;
    i128 fl = v >> s, rem = v - shl(fl, s), half = (i128)1 << (s - 1);
    switch (mode) {
    case 0: return fl + (rem >= half);
    case 1: return fl + (rem > half);
    case 2: return fl + (rem > half || (rem == half && v < 0));
    case 3: return fl + (rem > half || (rem == half && v > 0));
    case 4: return rem > half ? fl + 1 : (rem < half ? fl : fl + (fl & 1));
    case 6: return fl + (v < 0 && rem != 0);
    default: return fl;
    }
}
constexpr i128 overflow(i128 v, const Spec &t)
{
    const int W = 1 + t.i + t.f;
    const i128 mx = ((i128)1 << (W - 1)) - 1, mn = t.s ? -((i128)1 << (W - 1)) : (i128)0;
    switch (t.o) {
    case 0: return v > mx ? mx : (v < mn ? mn : v);
    case 1: return (v > mx || v < mn) ? (i128)0 : v;
    case 2: { const i128 lo = t.s ? -mx : (i128)0; return v > mx ? mx : (v < lo ? lo : v); }
    default:
        if (t.s) { const i128 m = ((i128)1 << W) - 1; i128 low = v & m; if (low >> (W - 1)) low |= ~m; return low; }
        return v & mx;
    }
}
constexpr i128 convert(i128 raw, int from_frac, const Spec &to) { return overflow(round_shift(raw, from_frac - to.f, to.q), to); }

// ---- tag parsing: tags in any order, all optional, first occurrence wins ----
struct Partial {
    bool has_i = false, has_f = false, has_s = false, has_q = false, has_o = false, full = false;
    int i = 0, f = 0; bool s = true; int q = 5, o = 0;
};
template <class T> struct tag { static constexpr void apply(Partial &) {} };                    // unknown tags are ignored
template <int V> struct tag<intBits<V>> { static constexpr void apply(Partial &p) { if (!p.has_i) { p.has_i = true; p.i = V; } } };
template <int V> struct tag<fracBits<V>> { static constexpr void apply(Partial &p) { if (!p.has_f) { p.has_f = true; p.f = V; } } };
template <bool V> struct tag<isSigned<V>> { static constexpr void apply(Partial &p) { if (!p.has_s) { p.has_s = true; p.s = V; } } };
template <class M> struct tag<QuMode<M>> { static constexpr void apply(Partial &p) { if (!p.has_q) { p.has_q = true; p.q = M::value; } } };
template <class M> struct tag<OfMode<M>> { static constexpr void apply(Partial &p) { if (!p.has_o) { p.has_o = true; p.o = M::value; } } };
template <> struct tag<FullPrec> { static constexpr void apply(Partial &p) { p.full = true; } };
template <class... Ts> struct tag<TypeList<Ts...>> { static constexpr void apply(Partial &p) { (tag<Ts>::apply(p), ...); } };
template <int I, int F, bool S, class Q, class O>
struct tag<Qu_s<intBits<I>, fracBits<F>, isSigned<S>, QuMode<Q>, OfMode<O>>> {       // a full type = its five tags
    static constexpr void apply(Partial &p) { tag<intBits<I>>::apply(p); tag<fracBits<F>>::apply(p); tag<isSigned<S>>::apply(p);
                                               tag<QuMode<Q>>::apply(p); tag<OfMode<O>>::apply(p); }
};
template <class... Ts> constexpr Partial parse() { Partial p; (tag<Ts>::apply(p), ...); return p; }

constexpr Spec merge(const Spec &a, const Spec &b, const Partial &t, bool is_mul)
{
    const int di = is_mul ? (t.full ? a.i + b.i : (a.i > b.i ? a.i : b.i)) : ((a.i > b.i ? a.i : b.i) + (t.full ? 1 : 0));
    const int df = is_mul && t.full ? a.f + b.f : (a.f > b.f ? a.f : b.f);
    return Spec{ t.has_i ? t.i : di, t.has_f ? t.f : df, t.has_s ? t.s : (a.s || b.s), t.has_q ? t.q : (a.q == b.q ? a.q : 5),
                 t.has_o ? t.o : (a.o == b.o ? a.o : 0) };
}

template <int V> struct qu_of_id;
template <> struct qu_of_id<0> { using type = RND::POS_INF; };
template <> struct qu_of_id<1> { using type = RND::NEG_INF; };
template <> struct qu_of_id<2> { using type = RND::ZERO; };
template <> struct qu_of_id<3> { using type = RND::INF; };
template <> struct qu_of_id<4> { using type = RND::CONV; };
template <> struct qu_of_id<5> { using type = TRN::TCPL; };
template <> struct qu_of_id<6> { using type = TRN::SMGN; };
template <int V> struct of_of_id;
template <> struct of_of_id<0> { using type = SAT::TCPL; };
template <> struct of_of_id<1> { using type = SAT::ZERO; };
template <> struct of_of_id<2> { using type = SAT::SMGN; };
template <> struct of_of_id<3> { using type = WRP::TCPL; };

template <Spec S>
using scalar_of = Qu_s<intBits<S.i>, fracBits<S.f>, isSigned<S.s>, QuMode<typename qu_of_id<S.q>::type>, OfMode<typename of_of_id<S.o>::type>>;

template <class T> struct is_scalar : std::false_type {};
template <int I, int F, bool S, class Q, class O>
struct is_scalar<Qu_s<intBits<I>, fracBits<F>, isSigned<S>, QuMode<Q>, OfMode<O>>> : std::true_type {};
template <class T> struct is_complex : std::false_type {};
template <class... A, class... B> struct is_complex<Qu_s<Qu_s<A...>, Qu_s<B...>>> : std::true_type {};
template <class T> struct is_dim : std::false_type {};
template <std::size_t... D> struct is_dim<dim<D...>> : std::true_type {};
template <class T> struct is_tensor : std::false_type {};
template <std::size_t... D, class E> struct is_tensor<Qu_s<dim<D...>, E>> : std::true_type {};
template <class T> struct is_typelist : std::false_type {};
template <class... Ts> struct is_typelist<TypeList<Ts...>> : std::true_type {};

// Qu<...> grammar
template <class... Args> struct make_qu {
    static constexpr Partial p = parse<Args...>();
    static constexpr Spec s{ p.has_i ? p.i : defaultIntBits, p.has_f ? p.f : defaultFracBits, p.has_s ? p.s : defaultIsSigned,
                             p.has_q ? p.q : defaultQuMode::value, p.has_o ? p.o : defaultOfMode::value };
    using type = scalar_of<s>;
};
template <class... A> struct make_qu<Qu_s<A...>> { using type = Qu_s<A...>; };
template <class... A, class... B> struct make_qu<Qu_s<A...>, Qu_s<B...>> { using type = Qu_s<Qu_s<A...>, Qu_s<B...>>; };
template <std::size_t... D, class... Rest> struct make_qu<dim<D...>, Rest...> { using type = Qu_s<dim<D...>, typename make_qu<Rest...>::type>; };

inline void check(int rc)
{
    if (rc != 0) throw std::runtime_error(std::string("QuBLAS (B200): ") + qb200_last_error());
}

inline std::mt19937 &rng() { static std::mt19937 g(1); return g; }          // reference seeds gen(1), QuBLAS.h:30

}  // namespace detail

template <class... Args> using Qu = typename detail::make_qu<Args...>::type;
template <class T> concept isScalar = detail::is_scalar<T>::value || detail::is_complex<T>::value;

// ------------------------------------------------------------------------------------------------
// raw word: the reference's ArbiInt<N>::data (int32 for N <= 32, int64 for N <= 64)
// ------------------------------------------------------------------------------------------------
template <std::size_t N>
    requires(N > 0 && N <= 64)
struct ArbiInt {
    using data_t = std::conditional_t<(N <= 32), std::int32_t, std::int64_t>;
    static constexpr std::size_t num_bits = N;
    data_t data = 0;
    constexpr ArbiInt() = default;
    constexpr ArbiInt(data_t v) : data(v) {}
    constexpr operator data_t() const { return data; }
    double toDouble() const { return (double)data; }
};

// ------------------------------------------------------------------------------------------------
// scalar fixed-point value
// ------------------------------------------------------------------------------------------------
template <int I, int F, bool S, class QM, class OM>
class Qu_s<intBits<I>, fracBits<F>, isSigned<S>, QuMode<QM>, OfMode<OM>> {
public:
    static_assert(0 <= I + F, "The total number of bits must be non-negative.");
    static_assert(1 + I + F <= 64, "types wider than 64 bits are outside the GPU parity domain (see DESIGN.md)");
    static constexpr int intB = I, fracB = F;
    static constexpr bool isS = S;
    static constexpr int width = I + F + (S ? 1 : 0);
    using QuM_t = QM; using OfM_t = OM;
    static constexpr int QuM = QM::value, OfM = OM::value;
    static constexpr bool is_complex = false;
    static constexpr detail::Spec spec{ I, F, S, QM::value, OM::value };
    using raw_t = typename ArbiInt<1 + I + F>::data_t;

    ArbiInt<1 + I + F> data;      // always has a sign bit, like the reference (QuBLAS.h:2384-2385)

    constexpr Qu_s() = default;

    // from a double: exact value, rounded with this type's QuMode, overflow by its OfMode (QuBLAS.h:2387-2393)
    constexpr Qu_s(double v)
    {
        if (!(v == v) || v == 0.0) { data.data = 0; return; }
        int e = 0;
        double m = std::frexp(v, &e);                       // v = m * 2^e, 0.5 <= |m| < 1
        const detail::i128 mi = (detail::i128)(long long)std::ldexp(m, 53);   // exact 53-bit mantissa
        int sh = 53 - e - F;                                // raw = mi * 2^(-sh)
        detail::i128 r;
        if (sh >= 0) r = detail::round_shift(mi, sh, QuM);
        else if (-sh > 70) r = mi > 0 ? ((detail::i128)1 << 124) : -((detail::i128)1 << 124);   // beyond every 64-bit type, low 64 bits zero
        else r = detail::shl(mi, -sh);
        data.data = (raw_t)detail::overflow(r, spec);
    }
    constexpr Qu_s(int v) : Qu_s((double)v) {}

    // converting ctor: the TARGET's modes (QuBLAS.h:2398-2411)
    template <int I2, int F2, bool S2, class Q2, class O2>
    constexpr Qu_s(const Qu_s<intBits<I2>, fracBits<F2>, isSigned<S2>, QuMode<Q2>, OfMode<O2>> &o)
    {
        data.data = (raw_t)detail::convert((detail::i128)o.data.data, F2, spec);
    }

    double toDouble() const { return std::ldexp((double)data.data, -F); }
    auto fill()
    {
        const long long lo = -(1LL << (I + F)), hi = (1LL << (I + F)) - 1;
        data.data = (raw_t)std::uniform_int_distribution<long long>(lo, hi)(detail::rng());
        return *this;
    }
    auto fill(int raw) { data.data = (raw_t)raw; return *this; }
    // the low `width` bits of the two's-complement raw word, MSB first (QuBLAS.h:2433-2438)
    std::string toString() const
    {
        std::string out((std::size_t)width, '0');
        const unsigned long long u = (unsigned long long)(long long)data.data;
        for (int b = 0; b < width; b++) if ((u >> b) & 1ULL) out[(std::size_t)(width - 1 - b)] = '1';
        return out;
    }
    void display(const std::string &name = "") const
    {
        if (!name.empty()) std::cout << name << " :\n";
        std::cout << "intBits: " << I << " fracBits: " << F << " isSigned: " << S << "\nraw: " << (long long)data.data
                  << "\nDecimal: " << toDouble() << "\n\n";
    }
    friend std::ostream &operator<<(std::ostream &os, const Qu_s &q) { return os << q.toDouble(); }
};

// ------------------------------------------------------------------------------------------------
// complex value {real; imag} (QuBLAS.h:2501-2617)
// ------------------------------------------------------------------------------------------------
template <class... A, class... B>
class Qu_s<Qu_s<A...>, Qu_s<B...>> {
public:
    using realType = Qu_s<A...>;
    using imagType = Qu_s<B...>;
    static constexpr bool is_complex = true;
    static constexpr int realWidth = realType::width, imagWidth = imagType::width, width = realWidth + imagWidth;   // :2507-2509
    realType real;
    imagType imag;
    constexpr Qu_s() = default;
    template <class T1, class T2> constexpr Qu_s(T1 a, T2 b) : real(a), imag(b) {}
    template <class T> requires(!detail::is_complex<T>::value) constexpr Qu_s(T a) : real(a), imag(0) {}
    template <class R2, class I2> constexpr Qu_s(const Qu_s<R2, I2> &o) requires(detail::is_scalar<R2>::value) : real(o.real), imag(o.imag) {}
    auto fill() { real.fill(); imag.fill(); return *this; }
    auto fill(int raw) { real.fill(raw); imag.fill(raw); return *this; }
    auto fill(int re_raw, int im_raw) { real.fill(re_raw); imag.fill(im_raw); return *this; }                        // :2571-2576
    std::complex<double> toDouble() const { return { real.toDouble(), imag.toDouble() }; }                            // :2548-2551
    friend std::ostream &operator<<(std::ostream &os, const Qu_s &q) { return os << q.toDouble(); }
    std::string toString() const { return "(" + real.toString() + ", " + imag.toString() + ")"; }                    // :2553-2556
    void display(const std::string &name = "") const { real.display(name + ".real"); imag.display(name + ".imag"); }
};
template <class R, class Im> using Qcomplex = Qu_s<R, Im>;

namespace detail {
// storage word of an element inside a tensor buffer handed to the C ABI
template <class E> struct elem_words;
template <int I, int F, bool S, class Q, class O>
struct elem_words<Qu_s<intBits<I>, fracBits<F>, isSigned<S>, QuMode<Q>, OfMode<O>>> {
    using E = Qu_s<intBits<I>, fracBits<F>, isSigned<S>, QuMode<Q>, OfMode<O>>;
    static_assert(sizeof(E) == sizeof(typename E::raw_t), "scalar must be exactly one raw word");
};
constexpr qb_qdesc lower(const Spec &s) { return qb_qdesc{ (int16_t)s.i, (int16_t)s.f, (uint8_t)s.s, (uint8_t)s.q, (uint8_t)s.o, 0 }; }
}  // namespace detail

// ------------------------------------------------------------------------------------------------
// tensor: column-major, one raw word per element, contiguous (QuBLAS.h:2675-3037)
// ------------------------------------------------------------------------------------------------
template <std::size_t... D, class E>
class Qu_s<dim<D...>, E> {
public:
    using size = dim<D...>;
    using elem_t = E;
    static constexpr std::size_t elemSize = size::elemSize, dimSize = size::dimSize;
    static constexpr bool onHeap = elemSize > 1000;                      // same storage rule as the reference (:2694-2697)
    using storage_t = std::conditional_t<onHeap, std::vector<E>, std::array<E, elemSize>>;
    storage_t data;

    Qu_s() { if constexpr (onHeap) data.resize(elemSize); }
    template <class... V>
        requires(sizeof...(V) > 0 && (std::is_convertible_v<V, E> && ...))
    Qu_s(V... values)
    {
        if constexpr (onHeap) data.resize(elemSize);
        if constexpr (sizeof...(V) == elemSize) { std::size_t k = 0; ((data[k++] = E(values)), ...); }
        else throw std::invalid_argument("The number of values does not match the size of the Qu_s.");
    }
    template <class E2>
    Qu_s(const Qu_s<dim<D...>, E2> &o)                                   // element-wise converting copy (:2759-2771)
    {
        if constexpr (onHeap) data.resize(elemSize);
        for (std::size_t i = 0; i < elemSize; i++) data[i] = E(o.data[i]);
    }
    template <class E2> Qu_s &operator=(const Qu_s<dim<D...>, E2> &o)
    {
        for (std::size_t i = 0; i < elemSize; i++) data[i] = E(o.data[i]);
        return *this;
    }

    static constexpr std::size_t linear(std::initializer_list<std::size_t> idx)
    {
        std::size_t acc = 0, mul = 1, k = 0;
        for (std::size_t v : idx) { acc += v * mul; mul *= size::dimArray[k++]; }
        return acc;
    }
    template <class... Ix>
        requires(sizeof...(Ix) == dimSize && sizeof...(Ix) > 1)
    E &operator[](Ix... ix) { return data[linear({ (std::size_t)ix... })]; }
    template <class... Ix>
        requires(sizeof...(Ix) == dimSize && sizeof...(Ix) > 1)
    const E &operator[](Ix... ix) const { return data[linear({ (std::size_t)ix... })]; }
    E &operator[](std::size_t i) { return data[i]; }
    const E &operator[](std::size_t i) const { return data[i]; }

    auto fill() { for (auto &e : data) e.fill(); return *this; }
    void clear() { std::memset((void *)data.data(), 0, elemSize * sizeof(E)); }
    // values as doubles, 4 decimals: "[a, b, ...]", a 2-D tensor as right-aligned rows (QuBLAS.h:2913-2977)
    friend std::ostream &operator<<(std::ostream &os, const Qu_s &t)
    {
        const auto flags = os.flags();
        const auto prec = os.precision();
        os << std::fixed << std::setprecision(4) << "[";
        if constexpr (dimSize == 2) {
            std::size_t cell = 0;
            for (std::size_t i = 0; i < elemSize; i++) {
                std::ostringstream probe;
                probe << std::fixed << std::setprecision(4) << t.data[i].toDouble();
                cell = std::max(cell, probe.str().size());
            }
            constexpr std::size_t rows = size::dimArray[0], cols = size::dimArray[1];
            for (std::size_t r = 0; r < rows; r++) {
                if (r) os << ' ';
                for (std::size_t c = 0; c < cols; c++) os << std::setw((int)cell) << std::right << t.data[r + c * rows].toDouble() << (c + 1 < cols ? ", " : "");
                if (r + 1 < rows) os << '\n';
            }
        } else {
            for (std::size_t i = 0; i < elemSize; i++) os << t.data[i] << (i + 1 < elemSize ? ", " : "");
        }
        os << "]";
        os.flags(flags);
        os.precision(prec);
        return os;
    }
    void *raw() { return (void *)data.data(); }                          // the buffer the C ABI sees
    const void *raw() const { return (const void *)data.data(); }
    void display(const std::string &name = "") const
    {
        if (!name.empty()) std::cout << name << " :\n";
        for (std::size_t i = 0; i < elemSize; i++) {
            if constexpr (detail::is_complex<E>::value) std::cout << "(" << data[i].real.toDouble() << "," << data[i].imag.toDouble() << ") ";
            else std::cout << data[i].toDouble() << " ";
        }
        std::cout << "\n";
    }
};

// ------------------------------------------------------------------------------------------------
// scalar operations (header-only; same result types as MulMerger / AddMerger, QuBLAS.h:3104-3140)
// ------------------------------------------------------------------------------------------------
template <class... Args> struct TFComplexMul {};
template <class... Args> struct BasicComplexMul {};
template <class... Args> struct acT {};
template <class... Args> struct bdT {};
template <class... Args> struct adT {};
template <class... Args> struct bcT {};
template <class... Args> struct acbdT {};
template <class... Args> struct adbcT {};
template <class... Args> struct abT {};
template <class... Args> struct cdT {};
template <class... Args> struct baT {};
template <class... Args> struct abcT {};
template <class... Args> struct cdbT {};
template <class... Args> struct badT {};
template <class... Args> struct ABT {};
template <class... Args> struct BCT {};

namespace detail {
template <class... toArgs, class A, class B> constexpr auto mul_s(const A &a, const B &b)
{
    constexpr Spec t = merge(A::spec, B::spec, parse<toArgs...>(), true);
    scalar_of<t> r;
    r.data.data = (typename scalar_of<t>::raw_t)convert((i128)a.data.data * (i128)b.data.data, A::fracB + B::fracB, t);   // :3157-3161
    return r;
}
template <bool Sub, class... toArgs, class A, class B> constexpr auto addsub_s(const A &a, const B &b)
{
    constexpr Spec t = merge(A::spec, B::spec, parse<toArgs...>(), false);
    constexpr int mf = A::fracB > B::fracB ? A::fracB : B::fracB;
    const i128 x = shl((i128)a.data.data, mf - A::fracB), y = shl((i128)b.data.data, mf - B::fracB);                      // :3182-3190
    scalar_of<t> r;
    r.data.data = (typename scalar_of<t>::raw_t)convert(Sub ? x - y : x + y, mf, t);
    return r;
}

// tag lookup inside a TFComplexMul argument pack: first match wins; a missing tag means "the whole
// pack" which only matters for FullPrec (reference: tagExtractor default, QuBLAS.h:3513-3520).
template <template <class...> class Tag, class... Pack> struct find_tag { static constexpr bool found = false; using list = TypeList<>; };
template <template <class...> class Tag, class... In, class... Rest>
struct find_tag<Tag, Tag<In...>, Rest...> { static constexpr bool found = true; using list = TypeList<In...>; };
template <template <class...> class Tag, class First, class... Rest>
struct find_tag<Tag, First, Rest...> : find_tag<Tag, Rest...> {};

template <std::size_t N, class... Ts> struct type_at_c;
template <std::size_t N, class H, class... T> struct type_at_c<N, H, T...> : type_at_c<N - 1, T...> {};
template <class H, class... T> struct type_at_c<0, H, T...> { using type = H; };

template <template <class...> class Tag, class... Pack> constexpr Partial tf_partial()
{
    using F = find_tag<Tag, Pack...>;
    if constexpr (F::found) return parse<typename F::list>();
    else { Partial p; p.full = (std::is_same_v<FullPrec, Pack> || ...); return p; }
}

struct TfSpec { Spec ab, cd, ba, abc, bad, cdb, AB, BC; };
// Resolved node types of TFComplexMul (QuBLAS.h:3510-3535): product B=(c+d)*b is governed by badT,
// C=(b-a)*d by cdbT, and baT is ignored (b-a gets the default AddMerger) — exactly as the reference.
template <class... Pack> constexpr TfSpec tf_resolve(const Spec &a, const Spec &b, const Spec &c, const Spec &d)
{
    TfSpec r{};
    r.ab = merge(a, b, tf_partial<abT, Pack...>(), false);
    r.cd = merge(c, d, tf_partial<cdT, Pack...>(), false);
    r.ba = merge(b, a, Partial{}, false);
    r.abc = merge(r.ab, c, tf_partial<abcT, Pack...>(), true);
    r.bad = merge(r.cd, b, tf_partial<badT, Pack...>(), true);
    r.cdb = merge(r.ba, d, tf_partial<cdbT, Pack...>(), true);
    r.AB = merge(r.abc, r.bad, tf_partial<ABT, Pack...>(), false);
    r.BC = merge(r.bad, r.cdb, tf_partial<BCT, Pack...>(), false);
    return r;
}
struct BcmSpec { Spec ac, bd, ad, bc, acbd, adbc; };
// Resolved node types of BasicComplexMul (QuBLAS.h:3426-3446): re = ac - bd, im = ad + bc.
template <class... Pack> constexpr BcmSpec bcm_resolve(const Spec &a, const Spec &b, const Spec &c, const Spec &d)
{
    BcmSpec r{};
    r.ac = merge(a, c, tf_partial<acT, Pack...>(), true);
    r.bd = merge(b, d, tf_partial<bdT, Pack...>(), true);
    r.ad = merge(a, d, tf_partial<adT, Pack...>(), true);
    r.bc = merge(b, c, tf_partial<bcT, Pack...>(), true);
    r.acbd = merge(r.ac, r.bd, tf_partial<acbdT, Pack...>(), false);
    r.adbc = merge(r.ad, r.bc, tf_partial<adbcT, Pack...>(), false);
    return r;
}
constexpr qb_bcm_types lower(const BcmSpec &t)
{
    return qb_bcm_types{ lower(t.ac), lower(t.bd), lower(t.ad), lower(t.bc), lower(t.acbd), lower(t.adbc) };
}
template <class T> struct bcm_pack { static constexpr bool is_bcm = false; };
template <class... P> struct bcm_pack<BasicComplexMul<P...>> {
    static constexpr bool is_bcm = true;
    static constexpr BcmSpec resolve(const Spec &a, const Spec &b, const Spec &c, const Spec &d) { return bcm_resolve<P...>(a, b, c, d); }
};
constexpr qb_tfc_types lower(const TfSpec &t)
{
    return qb_tfc_types{ lower(t.ab), lower(t.cd), lower(t.ba), lower(t.abc), lower(t.bad), lower(t.cdb), lower(t.AB), lower(t.BC) };
}
template <class T> struct tf_pack { static constexpr bool is_tf = false; };
template <class... P> struct tf_pack<TFComplexMul<P...>> {
    static constexpr bool is_tf = true;
    static constexpr TfSpec resolve(const Spec &a, const Spec &b, const Spec &c, const Spec &d) { return tf_resolve<P...>(a, b, c, d); }
};
}  // namespace detail

template <class... toArgs, class A, class B>
    requires(detail::is_scalar<A>::value && detail::is_scalar<B>::value)
constexpr auto Qmul(const A a, const B b) { return detail::mul_s<toArgs...>(a, b); }
template <class... toArgs, class A, class B>
    requires(detail::is_scalar<A>::value && detail::is_scalar<B>::value)
constexpr auto Qadd(const A a, const B b) { return detail::addsub_s<false, toArgs...>(a, b); }
template <class... toArgs, class A, class B>
    requires(detail::is_scalar<A>::value && detail::is_scalar<B>::value)
constexpr auto Qsub(const A a, const B b) { return detail::addsub_s<true, toArgs...>(a, b); }

// complex scalar product with TFComplexMul: 3 multipliers / 5 adders (QuBLAS.h:3522-3534)
template <class TF, class X, class Y>
    requires(detail::tf_pack<TF>::is_tf && detail::is_complex<X>::value && detail::is_complex<Y>::value)
constexpr auto Qmul(const X x, const Y y)
{
    constexpr detail::TfSpec t = detail::tf_pack<TF>::resolve(X::realType::spec, X::imagType::spec, Y::realType::spec, Y::imagType::spec);
    auto ab = Qadd<detail::scalar_of<t.ab>>(x.real, x.imag);
    auto cd = Qadd<detail::scalar_of<t.cd>>(y.real, y.imag);
    auto ba = Qsub<detail::scalar_of<t.ba>>(x.imag, x.real);
    auto A = Qmul<detail::scalar_of<t.abc>>(ab, y.real);
    auto B = Qmul<detail::scalar_of<t.bad>>(cd, x.imag);
    auto C = Qmul<detail::scalar_of<t.cdb>>(ba, y.imag);
    auto re = Qsub<detail::scalar_of<t.AB>>(A, B);
    auto im = Qsub<detail::scalar_of<t.BC>>(B, C);
    return Qcomplex<decltype(re), decltype(im)>(re, im);
}
// complex scalar product with BasicComplexMul: 4 multipliers / 2 adders (QuBLAS.h:3437-3445);
// Qmul(c1, c2) without a method uses it with default types (QuBLAS.h:3422-3424)
template <class BCM = BasicComplexMul<>, class X, class Y>
    requires(detail::bcm_pack<BCM>::is_bcm && detail::is_complex<X>::value && detail::is_complex<Y>::value)
constexpr auto Qmul(const X x, const Y y)
{
    constexpr detail::BcmSpec t = detail::bcm_pack<BCM>::resolve(X::realType::spec, X::imagType::spec, Y::realType::spec, Y::imagType::spec);
    auto ac = Qmul<detail::scalar_of<t.ac>>(x.real, y.real);
    auto bd = Qmul<detail::scalar_of<t.bd>>(x.imag, y.imag);
    auto ad = Qmul<detail::scalar_of<t.ad>>(x.real, y.imag);
    auto bc = Qmul<detail::scalar_of<t.bc>>(x.imag, y.real);
    auto re = Qsub<detail::scalar_of<t.acbd>>(ac, bd);
    auto im = Qadd<detail::scalar_of<t.adbc>>(ad, bc);
    return Qcomplex<decltype(re), decltype(im)>(re, im);
}
// realT<...> / imagT<...>: per-component target of a complex-valued result (QuBLAS.h:3537-3547).
// Without them the whole pack serves both components; two bare scalar types are positional
// (first = real part, second = imaginary part, :3567-3569).
template <class... Args> struct realT {};
template <class... Args> struct imagT {};
namespace detail {
template <bool Imag, class... Pack> constexpr Partial ri_partial()
{
    if constexpr (Imag ? find_tag<imagT, Pack...>::found : find_tag<realT, Pack...>::found) {
        if constexpr (Imag) return parse<typename find_tag<imagT, Pack...>::list>();
        else return parse<typename find_tag<realT, Pack...>::list>();
    } else if constexpr (sizeof...(Pack) == 2 && (is_scalar<Pack>::value && ...)) {
        return parse<typename type_at_c<Imag ? 1 : 0, Pack...>::type>();
    } else return parse<Pack...>();
}
template <bool Imag, class L> struct ri_partial_l;
template <bool Imag, class... Pack> struct ri_partial_l<Imag, TypeList<Pack...>> { static constexpr Partial value = ri_partial<Imag, Pack...>(); };
}  // namespace detail

// component-wise complex add / subtract (QuBLAS.h:3549-3590)
template <class... toArgs, class X, class Y>
    requires(detail::is_complex<X>::value && detail::is_complex<Y>::value)
constexpr auto Qadd(const X x, const Y y)
{
    constexpr detail::Spec rt = detail::merge(X::realType::spec, Y::realType::spec, detail::ri_partial<false, toArgs...>(), false);
    constexpr detail::Spec it = detail::merge(X::imagType::spec, Y::imagType::spec, detail::ri_partial<true, toArgs...>(), false);
    auto re = Qadd<detail::scalar_of<rt>>(x.real, y.real);
    auto im = Qadd<detail::scalar_of<it>>(x.imag, y.imag);
    return Qcomplex<decltype(re), decltype(im)>(re, im);
}
template <class... toArgs, class X, class Y>
    requires(detail::is_complex<X>::value && detail::is_complex<Y>::value)
constexpr auto Qsub(const X x, const Y y)
{
    constexpr detail::Spec rt = detail::merge(X::realType::spec, Y::realType::spec, detail::ri_partial<false, toArgs...>(), false);
    constexpr detail::Spec it = detail::merge(X::imagType::spec, Y::imagType::spec, detail::ri_partial<true, toArgs...>(), false);
    auto re = Qsub<detail::scalar_of<rt>>(x.real, y.real);
    auto im = Qsub<detail::scalar_of<it>>(x.imag, y.imag);
    return Qcomplex<decltype(re), decltype(im)>(re, im);
}

// real (x) complex: the real operand scales both components (QuBLAS.h:3600-3637)
template <class... toArgs, class X, class Y>
    requires(detail::is_scalar<X>::value && detail::is_complex<Y>::value)
constexpr auto Qmul(const X x, const Y y)
{
    constexpr detail::Spec rt = detail::merge(X::spec, Y::realType::spec, detail::ri_partial<false, toArgs...>(), true);
    constexpr detail::Spec it = detail::merge(X::spec, Y::imagType::spec, detail::ri_partial<true, toArgs...>(), true);
    auto re = Qmul<detail::scalar_of<rt>>(x, y.real);
    auto im = Qmul<detail::scalar_of<it>>(x, y.imag);
    return Qcomplex<decltype(re), decltype(im)>(re, im);
}
template <class... toArgs, class X, class Y>
    requires(detail::is_complex<X>::value && detail::is_scalar<Y>::value)
constexpr auto Qmul(const X x, const Y y)
{
    constexpr detail::Spec rt = detail::merge(X::realType::spec, Y::spec, detail::ri_partial<false, toArgs...>(), true);
    constexpr detail::Spec it = detail::merge(X::imagType::spec, Y::spec, detail::ri_partial<true, toArgs...>(), true);
    auto re = Qmul<detail::scalar_of<rt>>(x.real, y);
    auto im = Qmul<detail::scalar_of<it>>(x.imag, y);
    return Qcomplex<decltype(re), decltype(im)>(re, im);
}
// real (+/-) complex: only the real parts meet; the imaginary part passes through unchanged, except
// for real - complex, where it is Qsub<toArgs...>(0 of the real operand's type, imag) (QuBLAS.h:3646-3696)
template <class... toArgs, class X, class Y>
    requires(detail::is_scalar<X>::value && detail::is_complex<Y>::value)
constexpr auto Qadd(const X x, const Y y)
{
    auto re = Qadd<toArgs...>(x, y.real);
    return Qcomplex<decltype(re), typename Y::imagType>(re, y.imag);
}
template <class... toArgs, class X, class Y>
    requires(detail::is_complex<X>::value && detail::is_scalar<Y>::value)
constexpr auto Qadd(const X x, const Y y)
{
    auto re = Qadd<toArgs...>(x.real, y);
    return Qcomplex<decltype(re), typename X::imagType>(re, x.imag);
}
template <class... toArgs, class X, class Y>
    requires(detail::is_scalar<X>::value && detail::is_complex<Y>::value)
constexpr auto Qsub(const X x, const Y y)
{
    auto re = Qsub<toArgs...>(x, y.real);
    auto im = Qsub<toArgs...>(X(), y.imag);
    return Qcomplex<decltype(re), decltype(im)>(re, im);
}
template <class... toArgs, class X, class Y>
    requires(detail::is_complex<X>::value && detail::is_scalar<Y>::value)
constexpr auto Qsub(const X x, const Y y)
{
    auto re = Qsub<toArgs...>(x.real, y);
    return Qcomplex<decltype(re), typename X::imagType>(re, x.imag);
}

// ------------------------------------------------------------------------------------------------
// Qreduce
// ------------------------------------------------------------------------------------------------
namespace detail {
template <class... Args> struct level_args { using list = TypeList<Args...>; };                 // Qreduce<T0, T1, ...>
template <class... Args> struct level_args<TypeList<Args...>> { using list = TypeList<Args...>; }; // Qreduce<TypeList<...>>

template <std::size_t N, class List> struct type_at;
template <std::size_t N, class H, class... T> struct type_at<N, TypeList<H, T...>> : type_at<N - 1, TypeList<T...>> {};
template <class H, class... T> struct type_at<0, TypeList<H, T...>> { using type = H; };

// level l uses entry min(l, n-1); no entries -> default AddMerger (QuBLAS.h:4913-4921)
template <class List, std::size_t L> constexpr Partial level_partial()
{
    if constexpr (List::size == 0) return Partial{};
    else return parse<typename type_at<(L < List::size ? L : List::size - 1), List>::type>();
}
template <class List, std::size_t L> constexpr Spec level_spec(const Spec &a, const Spec &b) { return merge(a, b, level_partial<List, L>(), false); }

constexpr std::size_t ceil_log2(std::size_t n) { std::size_t l = 0; while (((std::size_t)1 << l) < n) l++; return l; }

// resolved level types of the vector form over `len` leaves of type `leaf`
template <class List, std::size_t Depth> struct level_table {
    template <std::size_t L> static constexpr void step(Spec &cur, qb_levels &out)
    {
        cur = level_spec<List, L>(cur, cur);
        out.lv[L] = lower(cur);
    }
    template <std::size_t... Ls> static constexpr Spec run(Spec leaf, qb_levels &out, std::index_sequence<Ls...>)
    {
        Spec cur = leaf;
        (step<Ls>(cur, out), ...);
        return cur;
    }
    static constexpr Spec build(Spec leaf, qb_levels &out)
    {
        out.n = (uint32_t)Depth;
        return run(leaf, out, std::make_index_sequence<Depth>{});
    }
};
template <class List, std::size_t Len> struct reduce_lowering {
    static constexpr std::size_t depth = ceil_log2(Len);
    static_assert(depth <= QB_MAX_LEVELS, "reduction too deep");
    static constexpr qb_levels levels(Spec leaf) { qb_levels L{}; level_table<List, depth>::build(leaf, L); return L; }
    static constexpr Spec out(Spec leaf) { qb_levels L{}; return level_table<List, depth>::build(leaf, L); }
};

// variadic form on host scalars (QuBLAS.h:4924-4957)
template <class List, std::size_t Layer, class T0> constexpr auto reduce_var(const T0 &x) { return x; }
template <class List, std::size_t Layer, class T0, class T1, class... Ts> constexpr auto reduce_var(const T0 &a, const T1 &b, const Ts &...rest);
template <class List, std::size_t Layer, class... Ts, std::size_t... Is>
constexpr auto reduce_pairs(std::index_sequence<Is...>, const std::tuple<Ts...> &t)
{
    using Lv = TypeList<>;
    (void)sizeof(Lv);
    if constexpr (List::size == 0)
        return reduce_var<List, Layer + 1>(Qadd<>(std::get<2 * Is>(t), std::get<2 * Is + 1>(t))...);
    else
        return reduce_var<List, Layer + 1>(
            Qadd<typename type_at<(Layer < List::size ? Layer : List::size - 1), List>::type>(std::get<2 * Is>(t), std::get<2 * Is + 1>(t))...);
}
template <class List, std::size_t Layer, class T0, class T1, class... Ts>
constexpr auto reduce_var(const T0 &a, const T1 &b, const Ts &...rest)
{
    constexpr std::size_t n = 2 + sizeof...(Ts);
    auto t = std::tie(a, b, rest...);
    auto r = reduce_pairs<List, Layer>(std::make_index_sequence<n / 2>{}, t);
    if constexpr (n % 2 == 1) {
        if constexpr (List::size == 0) return Qadd<>(r, std::get<n - 1>(t));
        else return Qadd<typename type_at<(Layer < List::size ? Layer : List::size - 1), List>::type>(r, std::get<n - 1>(t));
    } else return r;
}
}  // namespace detail

// Qreduce<Args...>(q0, q1, ...): variadic form, header-only (QuBLAS.h:4953-4957)
template <class... Args, class T0, class... Ts>
    requires(isScalar<T0> && (isScalar<Ts> && ...))
constexpr auto Qreduce(const T0 q0, const Ts... qs)
{
    return detail::reduce_var<typename detail::level_args<Args...>::list, 0>(q0, qs...);
}

// Qreduce<Args...>(tensor): vector form, runs on the GPU through qb200_reduce_batched (QuBLAS.h:4986-5001)
template <class... Args, std::size_t... D, class E>
    requires(detail::is_scalar<E>::value)
auto Qreduce(const Qu_s<dim<D...>, E> &t)
{
    using List = typename detail::level_args<Args...>::list;
    constexpr std::size_t len = dim<D...>::elemSize;
    using Low = detail::reduce_lowering<List, len>;
    constexpr detail::Spec out = Low::out(E::spec);
    static constexpr qb_reduce_plan plan = [] {
        qb_reduce_plan p{};
        p.elem = detail::lower(E::spec);
        p.levels = Low::levels(E::spec);
        p.len = len;
        p.tree_form = QB_TREE_VECTOR;
        return p;
    }();
    detail::scalar_of<out> r;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    detail::check(qb200_reduce_batched(&plan, t.raw(), &r.data.data, nullptr, 0, 1, &ex));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Qgemul (readme.md:80-87): tags in any order, all optional, out-parameter first
// ------------------------------------------------------------------------------------------------
template <class... Args> struct QgemulAddArgs {};
template <class... Args> struct QgemulMulArgs {};
template <bool V> struct QgemulTransposedA { static constexpr bool value = V; };
template <bool V> struct QgemulTransposedB { static constexpr bool value = V; };

namespace detail {
template <class... Tags> struct gemul_tags {
    using add = TypeList<>; using mul = TypeList<>;
    static constexpr bool ta = false, tb = false, has_add = false, has_mul = false, has_ta = false, has_tb = false;
};
template <class... In, class... Rest> struct gemul_tags<QgemulAddArgs<In...>, Rest...> : gemul_tags<Rest...> {
    using add = std::conditional_t<gemul_tags<Rest...>::has_add, typename gemul_tags<Rest...>::add, typename level_args<In...>::list>;
    static constexpr bool has_add = true;
};
template <class... In, class... Rest> struct gemul_tags<QgemulMulArgs<In...>, Rest...> : gemul_tags<Rest...> {
    using mul = std::conditional_t<gemul_tags<Rest...>::has_mul, typename gemul_tags<Rest...>::mul, TypeList<In...>>;
    static constexpr bool has_mul = true;
};
template <bool V, class... Rest> struct gemul_tags<QgemulTransposedA<V>, Rest...> : gemul_tags<Rest...> {
    static constexpr bool ta = gemul_tags<Rest...>::has_ta ? gemul_tags<Rest...>::ta : V;
    static constexpr bool has_ta = true;
};
template <bool V, class... Rest> struct gemul_tags<QgemulTransposedB<V>, Rest...> : gemul_tags<Rest...> {
    static constexpr bool tb = gemul_tags<Rest...>::has_tb ? gemul_tags<Rest...>::tb : V;
    static constexpr bool has_tb = true;
};

template <class L> struct first_is_tf { static constexpr bool value = false; using type = void; };
template <class... P, class... R> struct first_is_tf<TypeList<TFComplexMul<P...>, R...>> { static constexpr bool value = true; using type = TFComplexMul<P...>; };

template <class L> struct first_is_bcm { using type = BasicComplexMul<>; };
template <class... P, class... R> struct first_is_bcm<TypeList<BasicComplexMul<P...>, R...>> { using type = BasicComplexMul<P...>; };
template <class TFL> struct tf_or_default {
    static constexpr TfSpec resolve(const Spec &a, const Spec &b, const Spec &c, const Spec &d)
    {
        if constexpr (TFL::value) return tf_pack<typename TFL::type>::resolve(a, b, c, d);
        else return tf_resolve<>(a, b, c, d);
    }
};

template <class T> struct dims_of;
template <std::size_t R, std::size_t C, class E> struct dims_of<Qu_s<dim<R, C>, E>> { static constexpr std::size_t rows = R, cols = C; using elem = E; };
template <std::size_t R, class E> struct dims_of<Qu_s<dim<R>, E>> { static constexpr std::size_t rows = R, cols = 1; using elem = E; };   // vector = R x 1
}  // namespace detail

template <class... Tags, class CT, class AT, class BT>
    requires(detail::is_tensor<CT>::value && detail::is_tensor<AT>::value && detail::is_tensor<BT>::value)
void Qgemul(CT &C, const AT &A, const BT &B)
{
    using G = detail::gemul_tags<Tags...>;
    using DA = detail::dims_of<AT>; using DB = detail::dims_of<BT>; using DC = detail::dims_of<CT>;
    constexpr std::size_t M = G::ta ? DA::cols : DA::rows, K = G::ta ? DA::rows : DA::cols;
    constexpr std::size_t Kb = G::tb ? DB::cols : DB::rows, N = G::tb ? DB::rows : DB::cols;
    static_assert(K == Kb, "Qgemul: inner dimensions do not match");
    static_assert(DC::rows == M && DC::cols == N, "Qgemul: output shape does not match");
    using EA = typename DA::elem; using EB = typename DB::elem; using EC = typename DC::elem;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    if constexpr (detail::is_scalar<EA>::value != detail::is_scalar<EB>::value) {
        // one real and one complex operand: real (x) complex products (QuBLAS.h:3600-3637), QgemulMulArgs<realT<..>, imagT<..>>
        // (or one type for both components), accumulated by the component-wise complex tree
        static_assert(detail::is_complex<EC>::value, "Qgemul: a real-by-complex product needs a complex output tensor");
        static constexpr bool real_a = detail::is_scalar<EA>::value;
        using ER = std::conditional_t<real_a, EA, EB>;                        // the real operand's element
        using EZ = std::conditional_t<real_a, EB, EA>;                        // the complex operand's element
        static constexpr detail::Spec leaf_re = detail::merge(real_a ? ER::spec : EZ::realType::spec, real_a ? EZ::realType::spec : ER::spec,
                                                              detail::ri_partial_l<false, typename G::mul>::value, true);
        static constexpr detail::Spec leaf_im = detail::merge(real_a ? ER::spec : EZ::imagType::spec, real_a ? EZ::imagType::spec : ER::spec,
                                                              detail::ri_partial_l<true, typename G::mul>::value, true);
        using Low = detail::reduce_lowering<typename G::add, K>;
        static constexpr qb_cgemul_plan plan = [] {
            qb_cgemul_plan p{};
            if constexpr (real_a) {
                p.a_re = p.a_im = detail::lower(ER::spec);
                p.b_re = detail::lower(EZ::realType::spec); p.b_im = detail::lower(EZ::imagType::spec);
                p.product = QB_CPROD_REAL_A; p.bcm.ac = detail::lower(leaf_re); p.bcm.ad = detail::lower(leaf_im);
            } else {
                p.a_re = detail::lower(EZ::realType::spec); p.a_im = detail::lower(EZ::imagType::spec);
                p.b_re = p.b_im = detail::lower(ER::spec);
                p.product = QB_CPROD_REAL_B; p.bcm.ac = detail::lower(leaf_re); p.bcm.bc = detail::lower(leaf_im);
            }
            p.add_re = Low::levels(leaf_re); p.add_im = Low::levels(leaf_im);
            p.c_re = detail::lower(EC::realType::spec); p.c_im = detail::lower(EC::imagType::spec);
            p.M = (uint32_t)M; p.N = (uint32_t)N; p.K = (uint32_t)K;
            p.trans_a = G::ta; p.trans_b = G::tb; p.tree_form = QB_TREE_VARIADIC;
            p.stride_a = (int64_t)(M * K); p.stride_b = (int64_t)(K * N); p.stride_c = (int64_t)(M * N);
            return p;
        }();
        detail::check(qb200_cgemul_batched(&plan, A.raw(), B.raw(), C.raw(), 1, &ex));
    } else if constexpr (detail::is_scalar<EA>::value) {
        static_assert(detail::is_scalar<EC>::value, "Qgemul: real operands need a real output tensor");
        static constexpr detail::Spec mul = detail::merge(EA::spec, EB::spec, detail::parse<typename G::mul>(), true);
        using Low = detail::reduce_lowering<typename G::add, K>;
        static constexpr qb_gemul_plan plan = [] {
            qb_gemul_plan p{};
            p.a = detail::lower(EA::spec); p.b = detail::lower(EB::spec); p.mul = detail::lower(mul); p.c = detail::lower(EC::spec);
            p.add = Low::levels(mul);
            p.M = (uint32_t)M; p.N = (uint32_t)N; p.K = (uint32_t)K;
            p.trans_a = G::ta; p.trans_b = G::tb; p.tree_form = QB_TREE_VARIADIC;
            p.stride_a = (int64_t)(M * K); p.stride_b = (int64_t)(K * N); p.stride_c = (int64_t)(M * N);
            return p;
        }();
        detail::check(qb200_gemul_batched(&plan, A.raw(), B.raw(), C.raw(), 1, &ex));
    } else {
        // complex elements: QgemulMulArgs<TFComplexMul<...>> (3 mult / 5 add) or BasicComplexMul<...> / nothing (4 mult / 2 add)
        using TFL = detail::first_is_tf<typename G::mul>;
        using BCL = detail::first_is_bcm<typename G::mul>;
        static constexpr detail::TfSpec tf = detail::tf_or_default<TFL>::resolve(EA::realType::spec, EA::imagType::spec, EB::realType::spec, EB::imagType::spec);
        static constexpr detail::BcmSpec bcm = detail::bcm_pack<typename BCL::type>::resolve(EA::realType::spec, EA::imagType::spec, EB::realType::spec, EB::imagType::spec);
        static constexpr detail::Spec leaf_re = TFL::value ? tf.AB : bcm.acbd, leaf_im = TFL::value ? tf.BC : bcm.adbc;
        using Low = detail::reduce_lowering<typename G::add, K>;
        static constexpr qb_cgemul_plan plan = [] {
            qb_cgemul_plan p{};
            p.a_re = detail::lower(EA::realType::spec); p.a_im = detail::lower(EA::imagType::spec);
            p.b_re = detail::lower(EB::realType::spec); p.b_im = detail::lower(EB::imagType::spec);
            if constexpr (TFL::value) { p.tf = detail::lower(tf); p.product = QB_CPROD_TF; }
            else { p.bcm = detail::lower(bcm); p.product = QB_CPROD_BASIC; }
            p.add_re = Low::levels(leaf_re); p.add_im = Low::levels(leaf_im);
            p.c_re = detail::lower(EC::realType::spec); p.c_im = detail::lower(EC::imagType::spec);
            p.M = (uint32_t)M; p.N = (uint32_t)N; p.K = (uint32_t)K;
            p.trans_a = G::ta; p.trans_b = G::tb; p.tree_form = QB_TREE_VARIADIC;
            p.stride_a = (int64_t)(M * K); p.stride_b = (int64_t)(K * N); p.stride_c = (int64_t)(M * N);
            return p;
        }();
        detail::check(qb200_cgemul_batched(&plan, A.raw(), B.raw(), C.raw(), 1, &ex));
    }
}

// ------------------------------------------------------------------------------------------------
// tensor Qmul<TFComplexMul<...>>(x, y): element-wise complex product on the GPU (readme.md:101-109)
// ------------------------------------------------------------------------------------------------
template <class TF, std::size_t... D, class EX, class EY>
    requires(detail::tf_pack<TF>::is_tf && detail::is_complex<EX>::value && detail::is_complex<EY>::value)
auto Qmul(const Qu_s<dim<D...>, EX> &x, const Qu_s<dim<D...>, EY> &y)
{
    static constexpr detail::TfSpec t = detail::tf_pack<TF>::resolve(EX::realType::spec, EX::imagType::spec, EY::realType::spec, EY::imagType::spec);
    using R = Qcomplex<detail::scalar_of<t.AB>, detail::scalar_of<t.BC>>;
    static_assert(sizeof(typename EX::realType) == sizeof(typename EX::imagType) && sizeof(typename EY::realType) == sizeof(typename EY::imagType) &&
                  sizeof(detail::scalar_of<t.AB>) == sizeof(detail::scalar_of<t.BC>),
                  "complex parts with different storage words are not supported on the GPU path");
    static constexpr qb_tfcmul_plan plan = [] {
        qb_tfcmul_plan p{};
        p.x_re = detail::lower(EX::realType::spec); p.x_im = detail::lower(EX::imagType::spec);
        p.y_re = detail::lower(EY::realType::spec); p.y_im = detail::lower(EY::imagType::spec);
        p.tf = detail::lower(t);
        p.len = dim<D...>::elemSize;
        return p;
    }();
    Qu_s<dim<D...>, R> out;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    detail::check(qb200_tfcmul_batched(&plan, x.raw(), y.raw(), out.raw(), 1, &ex));
    return out;
}

// tensor Qmul<BasicComplexMul<...>>(x, y) / Qmul(x, y): element-wise 4-mult/2-add complex product on the GPU
template <class BCM = BasicComplexMul<>, std::size_t... D, class EX, class EY>
    requires(detail::bcm_pack<BCM>::is_bcm && detail::is_complex<EX>::value && detail::is_complex<EY>::value)
auto Qmul(const Qu_s<dim<D...>, EX> &x, const Qu_s<dim<D...>, EY> &y)
{
    static constexpr detail::BcmSpec t = detail::bcm_pack<BCM>::resolve(EX::realType::spec, EX::imagType::spec, EY::realType::spec, EY::imagType::spec);
    using R = Qcomplex<detail::scalar_of<t.acbd>, detail::scalar_of<t.adbc>>;
    static_assert(sizeof(typename EX::realType) == sizeof(typename EX::imagType) && sizeof(typename EY::realType) == sizeof(typename EY::imagType) &&
                  sizeof(detail::scalar_of<t.acbd>) == sizeof(detail::scalar_of<t.adbc>),
                  "complex parts with different storage words are not supported on the GPU path");
    static constexpr qb_tfcmul_plan plan = [] {
        qb_tfcmul_plan p{};
        p.x_re = detail::lower(EX::realType::spec); p.x_im = detail::lower(EX::imagType::spec);
        p.y_re = detail::lower(EY::realType::spec); p.y_im = detail::lower(EY::imagType::spec);
        p.bcm = detail::lower(t);
        p.product = QB_CPROD_BASIC;
        p.len = dim<D...>::elemSize;
        return p;
    }();
    Qu_s<dim<D...>, R> out;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    detail::check(qb200_tfcmul_batched(&plan, x.raw(), y.raw(), out.raw(), 1, &ex));
    return out;
}

// tensor Qmul<realT<..>, imagT<..>>(real tensor, complex tensor) and the mirrored order: element-wise
// real (x) complex products (QuBLAS.h:3600-3637) on the GPU
namespace detail {
template <bool RealFirst, class... toArgs, std::size_t... D, class ER, class EZ>
auto tensor_rc(const void *x, const void *y, const Qu_s<dim<D...>, ER> *, const Qu_s<dim<D...>, EZ> *)
{
    static constexpr Spec tre = merge(RealFirst ? ER::spec : EZ::realType::spec, RealFirst ? EZ::realType::spec : ER::spec, ri_partial<false, toArgs...>(), true);
    static constexpr Spec tim = merge(RealFirst ? ER::spec : EZ::imagType::spec, RealFirst ? EZ::imagType::spec : ER::spec, ri_partial<true, toArgs...>(), true);
    using R = Qcomplex<scalar_of<tre>, scalar_of<tim>>;
    static_assert(sizeof(typename EZ::realType) == sizeof(typename EZ::imagType) && sizeof(scalar_of<tre>) == sizeof(scalar_of<tim>),
                  "complex parts with different storage words are not supported on the GPU path");
    static constexpr qb_tfcmul_plan plan = [] {
        qb_tfcmul_plan p{};
        if constexpr (RealFirst) {
            p.x_re = p.x_im = lower(ER::spec); p.y_re = lower(EZ::realType::spec); p.y_im = lower(EZ::imagType::spec);
            p.product = QB_CPROD_REAL_A; p.bcm.ac = lower(tre); p.bcm.ad = lower(tim);
        } else {
            p.x_re = lower(EZ::realType::spec); p.x_im = lower(EZ::imagType::spec); p.y_re = p.y_im = lower(ER::spec);
            p.product = QB_CPROD_REAL_B; p.bcm.ac = lower(tre); p.bcm.bc = lower(tim);
        }
        p.len = dim<D...>::elemSize;
        return p;
    }();
    Qu_s<dim<D...>, R> out;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    check(qb200_tfcmul_batched(&plan, x, y, out.raw(), 1, &ex));
    return out;
}
}  // namespace detail
template <class... toArgs, std::size_t... D, class EX, class EY>
    requires(detail::is_scalar<EX>::value && detail::is_complex<EY>::value)
auto Qmul(const Qu_s<dim<D...>, EX> &x, const Qu_s<dim<D...>, EY> &y) { return detail::tensor_rc<true, toArgs...>(x.raw(), y.raw(), &x, &y); }
template <class... toArgs, std::size_t... D, class EX, class EY>
    requires(detail::is_complex<EX>::value && detail::is_scalar<EY>::value)
auto Qmul(const Qu_s<dim<D...>, EX> &x, const Qu_s<dim<D...>, EY> &y) { return detail::tensor_rc<false, toArgs...>(x.raw(), y.raw(), &y, &x); }

// ------------------------------------------------------------------------------------------------
// element-wise Qmul / Qadd / Qsub of two real tensors: the reference returns lazy Mul/Add/SubExpression
// objects evaluated per element (QuBLAS.h:3741-3811, 4080-4156); here the same per-element result is
// computed eagerly on the GPU and returned as a tensor of the merger's result type.
// ------------------------------------------------------------------------------------------------
namespace detail {
template <int OP, class... toArgs, std::size_t... D, class EA, class EB>
auto tensor_ew(const Qu_s<dim<D...>, EA> &a, const Qu_s<dim<D...>, EB> &b)
{
    static constexpr Spec t = merge(EA::spec, EB::spec, parse<toArgs...>(), OP == QB_EW_MUL);
    static constexpr qb_ew_plan plan = [] {
        qb_ew_plan p{};
        p.a = lower(EA::spec); p.b = lower(EB::spec); p.to = lower(t);
        p.len = dim<D...>::elemSize; p.op = OP;
        return p;
    }();
    Qu_s<dim<D...>, scalar_of<t>> out;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    check(qb200_elementwise_batched(&plan, a.raw(), b.raw(), out.raw(), 1, &ex));
    return out;
}
}  // namespace detail
template <class... toArgs, std::size_t... D, class EA, class EB>
    requires(detail::is_scalar<EA>::value && detail::is_scalar<EB>::value)
auto Qmul(const Qu_s<dim<D...>, EA> &a, const Qu_s<dim<D...>, EB> &b) { return detail::tensor_ew<QB_EW_MUL, toArgs...>(a, b); }
template <class... toArgs, std::size_t... D, class EA, class EB>
    requires(detail::is_scalar<EA>::value && detail::is_scalar<EB>::value)
auto Qadd(const Qu_s<dim<D...>, EA> &a, const Qu_s<dim<D...>, EB> &b) { return detail::tensor_ew<QB_EW_ADD, toArgs...>(a, b); }
template <class... toArgs, std::size_t... D, class EA, class EB>
    requires(detail::is_scalar<EA>::value && detail::is_scalar<EB>::value)
auto Qsub(const Qu_s<dim<D...>, EA> &a, const Qu_s<dim<D...>, EB> &b) { return detail::tensor_ew<QB_EW_SUB, toArgs...>(a, b); }

// ------------------------------------------------------------------------------------------------
// ANUS: polynomial / piecewise-polynomial non-linearities (QuBLAS.h:4829-4897) and lookup tables (readme.md:66-78).
// Qpoly<a0, a1, ...>(x) evaluates a0 + x*(a1 + x*(...)) by Horner's rule; the product and the sum
// of each step are quantized to the type of that step's own coefficient (decltype(a0)).
// Qapprox<Segment<breakpoint, coeffs...>, ...>(x) picks the first segment whose breakpoint exceeds
// x.toDouble() (the last one otherwise) and converts the value back to x's type.
// On a scalar both run on the host, as in the reference.  On a TENSOR (no upstream counterpart: the reference's take one
// scalar) every element is evaluated on the GPU through qb200_poly_batched -- same quantizers, same bits as the scalar form.
// Qtable<func>(x) = decltype(x){ func(x.toDouble()) }; the reference documents it in its readme only, the three
// predefined tables are rsqrtFunc, reciprocalFunc and sqrtFunc; a tensor is looked up on the GPU (qb200_table_batched).
// ------------------------------------------------------------------------------------------------
namespace ANUS {
namespace detail {
template <auto... coeffs> constexpr qb_poly_segment lower_segment(double breakpoint)
{
    qb_poly_segment s{};
    s.breakpoint = breakpoint;
    s.ncoef = sizeof...(coeffs);
    std::size_t i = 0;
    ((s.coef[i].type = ::QuBLAS::detail::lower(std::remove_cvref_t<decltype(coeffs)>::spec), s.coef[i].raw = (int64_t)coeffs.data.data, ++i), ...);
    return s;
}
}  // namespace detail
template <auto a0> constexpr auto Qpoly(::QuBLAS::isScalar auto) { return a0; }
template <auto a0, auto a1, auto... more> constexpr auto Qpoly(::QuBLAS::isScalar auto x)
{
    return Qadd<decltype(a0)>(a0, Qmul<decltype(a0)>(x, Qpoly<a1, more...>(x)));
}
template <auto a0, auto... more, std::size_t... D, class E>
    requires(::QuBLAS::detail::is_scalar<E>::value && sizeof...(more) < QB_POLY_MAX_COEF)
auto Qpoly(const Qu_s<dim<D...>, E> &x)
{
    static constexpr qb_poly_plan plan = [] {
        qb_poly_plan p{};
        p.x = ::QuBLAS::detail::lower(E::spec); p.nseg = 1; p.approx = 0; p.len = dim<D...>::elemSize;
        p.seg[0] = detail::lower_segment<a0, more...>(0.0);
        return p;
    }();
    Qu_s<dim<D...>, std::remove_cvref_t<decltype(a0)>> out;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    ::QuBLAS::detail::check(qb200_poly_batched(&plan, x.raw(), out.raw(), 1, &ex));
    return out;
}
template <double BreakPoint, auto... coeffs> struct Segment {
    static constexpr double breakpoint = BreakPoint;
    static constexpr auto func(auto x) { return Qpoly<coeffs...>(x); }
    static constexpr qb_poly_segment lowered() { return detail::lower_segment<coeffs...>(BreakPoint); }
};
template <class First, class... Rest> constexpr auto Qapprox(::QuBLAS::isScalar auto x)
{
    if constexpr (sizeof...(Rest) == 0) return decltype(x){ First::func(x) };
    else {
        if (x.toDouble() < First::breakpoint) return decltype(x){ First::func(x) };
        return Qapprox<Rest...>(x);
    }
}
template <class... Segments, std::size_t... D, class E>
    requires(::QuBLAS::detail::is_scalar<E>::value && sizeof...(Segments) >= 1 && sizeof...(Segments) <= QB_POLY_MAX_SEG)
auto Qapprox(const Qu_s<dim<D...>, E> &x)
{
    static constexpr qb_poly_plan plan = [] {
        qb_poly_plan p{};
        p.x = ::QuBLAS::detail::lower(E::spec); p.nseg = sizeof...(Segments); p.approx = 1; p.len = dim<D...>::elemSize;
        std::size_t i = 0;
        ((p.seg[i++] = Segments::lowered()), ...);
        return p;
    }();
    Qu_s<dim<D...>, E> out;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    ::QuBLAS::detail::check(qb200_poly_batched(&plan, x.raw(), out.raw(), 1, &ex));
    return out;
}
template <class... Segments> struct Qapprox_s {
    static constexpr auto execute(const auto &x) { return Qapprox<Segments...>(x); }
};

inline constexpr auto rsqrtFunc = [](double v) { return 1.0 / std::sqrt(v); };
inline constexpr auto reciprocalFunc = [](double v) { return 1.0 / v; };
inline constexpr auto sqrtFunc = [](double v) { return std::sqrt(v); };
namespace detail {
template <auto func> double table_value(double v) { const double r = func(v); return std::isfinite(r) ? r : 0.0; }    // poles and domain errors read as 0
}
template <auto func> auto Qtable(::QuBLAS::isScalar auto x) { return decltype(x){ detail::table_value<func>(x.toDouble()) }; }
template <auto func, std::size_t... D, class E>
    requires(::QuBLAS::detail::is_scalar<E>::value && E::intB + E::fracB + 1 <= 20)
auto Qtable(const Qu_s<dim<D...>, E> &x)
{
    // one entry per raw value of E, built once per (func, E) with the scalar form
    static const std::vector<typename E::raw_t> table = [] {
        constexpr long long lo = -(1ll << (E::intB + E::fracB)), n = 1ll << (E::intB + E::fracB + 1);
        std::vector<typename E::raw_t> t((std::size_t)n);
        for (long long r = 0; r < n; r++) { E e; e.data.data = (typename E::raw_t)(lo + r); t[(std::size_t)r] = Qtable<func>(e).data.data; }
        return t;
    }();
    static constexpr qb_qdesc d = ::QuBLAS::detail::lower(E::spec);
    Qu_s<dim<D...>, E> out;
    qb_exec ex{ QB_MEM_HOST, -1, nullptr, 0, 0 };
    ::QuBLAS::detail::check(qb200_table_batched(&d, &d, table.data(), x.raw(), out.raw(), dim<D...>::elemSize, &ex));
    return out;
}
}  // namespace ANUS

// ------------------------------------------------------------------------------------------------
// BitStream: binary-string import/export of scalars and tensors — the test-vector format fed to HDL
// testbenches (QuBLAS.h:4545-4827, readme.md:111-127).  Host-only.
//   BitStream<elemOrder>(scalar)                 -> std::string         BitStream<ScalarT, elemOrder>(str) -> ScalarT
//   BitStream<tensorOrder, elemOrder>(tensor)    -> std::string         BitStream<TensorT, tensorOrder, elemOrder>(str) -> TensorT
// l2r keeps the order; r2l<n> reverses it in groups of n (characters of one element, or elements of
// the tensor).  Parsing ignores every character that is not '0'/'1' and stores the bits as the RAW
// word without sign extension, exactly like the reference (std::stoi(..., 2) + fill(int), :4697-4700).
// ------------------------------------------------------------------------------------------------
struct l2r {};
template <std::size_t... n>
    requires(sizeof...(n) <= 1)
struct r2l;
template <std::size_t n> struct r2l<n> { static constexpr std::size_t index = n; };
template <> struct r2l<> : r2l<1> {};

namespace detail {
template <class P> struct bit_order;
template <> struct bit_order<l2r> {
    static constexpr std::size_t group = 0;
    static std::string apply(std::string_view s) { return std::string(s); }
};
template <std::size_t... n> struct bit_order<r2l<n...>> {
    static constexpr std::size_t group = r2l<n...>::index;
    static std::string apply(std::string_view s)
    {
        if (s.size() % group != 0) throw std::runtime_error("Invalid string length: Must be a multiple of " + std::to_string(group));
        std::string out;
        out.reserve(s.size());
        for (std::size_t pos = s.size(); pos >= group; pos -= group) out.append(s.substr(pos - group, group));
        return out;
    }
};
// position in the serialized stream -> element index (the same map serves both directions)
template <class P> constexpr std::size_t stream_slot(std::size_t pos, std::size_t count)
{
    constexpr std::size_t g = bit_order<P>::group;
    if constexpr (g == 0) return pos;
    else return count - g * (pos / g + 1) + pos % g;
}
inline std::string only_bits(std::string_view s)
{
    std::string out;
    out.reserve(s.size());
    for (char ch : s) if (ch == '0' || ch == '1') out.push_back(ch);
    return out;
}
template <class E> E from_bits(const std::string &bits)
{
    E e;
    if constexpr (is_complex<E>::value)
        e.fill(std::stoi(bits.substr(0, (std::size_t)E::realWidth), nullptr, 2), std::stoi(bits.substr((std::size_t)E::realWidth, (std::size_t)E::imagWidth), nullptr, 2));
    else
        e.fill(std::stoi(bits, nullptr, 2));
    return e;
}
template <class... Args> struct bitstream;
template <class P> struct bitstream<P> {                                              // scalar -> string (:4787-4794)
    static std::string run(const auto &q) { return bit_order<P>::apply(q.toString()); }
};
template <class... QA, class P>
    requires(is_scalar<Qu_s<QA...>>::value || is_complex<Qu_s<QA...>>::value)
struct bitstream<Qu_s<QA...>, P> {                                                    // string -> scalar (:4760-4785)
    static auto run(std::string_view s) { return from_bits<Qu_s<QA...>>(bit_order<P>::apply(only_bits(s))); }
};
template <class TP, class EP> struct bitstream<TP, EP> {                              // tensor -> string (:4812-4821)
    template <std::size_t... D, class E> static std::string run(const Qu_s<dim<D...>, E> &t)
    {
        constexpr std::size_t n = dim<D...>::elemSize;
        static_assert(bit_order<TP>::group == 0 || n % bit_order<TP>::group == 0, "r2l<n> over the tensor needs elemSize % n == 0");
        std::string out;
        for (std::size_t pos = 0; pos < n; pos++) out.append(bit_order<EP>::apply(t[stream_slot<TP>(pos, n)].toString()));
        return out;
    }
};
template <std::size_t... D, class E, class TP, class EP>
struct bitstream<Qu_s<dim<D...>, E>, TP, EP> {                                        // string -> tensor (:4797-4810)
    static auto run(std::string_view s)
    {
        constexpr std::size_t n = dim<D...>::elemSize, w = (std::size_t)E::width;
        static_assert(bit_order<TP>::group == 0 || n % bit_order<TP>::group == 0, "r2l<n> over the tensor needs elemSize % n == 0");
        const std::string bits = only_bits(s);
        if (bits.size() % w != 0) throw std::runtime_error("Invalid string length: Must be a multiple of " + std::to_string(w));
        if (bits.size() < n * w) throw std::out_of_range("BitStream: the string holds fewer elements than the tensor");
        Qu_s<dim<D...>, E> t;
        for (std::size_t i = 0; i < n; i++) t[i] = from_bits<E>(bit_order<EP>::apply(std::string_view(bits).substr(stream_slot<TP>(i, n) * w, w)));
        return t;
    }
};
}  // namespace detail

template <class... Args> auto BitStream(const auto &input) { return detail::bitstream<Args...>::run(input); }

// ------------------------------------------------------------------------------------------------
// Batched ("Monte-Carlo") containers: `count` tensors of one type in ONE contiguous raw buffer, so a
// whole accuracy sweep is a single C-ABI call sharded by batch (tensors above 1000 elements are
// std::vector-backed in the reference, QuBLAS.h:2694-2697, so an array of tensors is not contiguous).
//
//  * Storage word: the NARROWEST two's-complement word that holds the element (int8 for stored width <= 8, int16 for
//    <= 16, int32, int64) -- the layout of this container is ours, and the host link is what bounds a batched call from
//    host memory: 16-bit operands cross PCIe as 2 bytes, not as the reference's 4.  Qwords::reference selects the
//    reference's ArbiInt word (int32 / int64) for zero-copy exchange with buffers in the reference's layout.
//  * Memory: Qmem::pinned (default; page-locked host memory from qb200_host_alloc, so the library's H2D / kernel / D2H
//    pipeline really overlaps), Qmem::pageable, or Qmem::device (the batch lives in HBM: synth(), QgemulBatched() and
//    QgemulError() then never cross PCIe -- a Monte-Carlo sweep returns 32 bytes of statistics).
// ------------------------------------------------------------------------------------------------
enum class Qmem { pinned, pageable, device, write_combined /* pinned, uncached on the CPU side: for input batches that the CPU only writes */ };
enum class Qwords { narrow, reference };

namespace detail {
template <int W, Qwords P> struct batch_word {
    using type = std::conditional_t<P == Qwords::reference, std::conditional_t<(W <= 32), std::int32_t, std::int64_t>,
                 std::conditional_t<(W <= 8), std::int8_t, std::conditional_t<(W <= 16), std::int16_t,
                 std::conditional_t<(W <= 32), std::int32_t, std::int64_t>>>>;
};
}  // namespace detail

template <class TensorT, Qwords Words = Qwords::narrow> class Qbatch {
public:
    using tensor_t = TensorT;
    using elem_t = typename TensorT::elem_t;
    static_assert(detail::is_scalar<elem_t>::value, "Qbatch holds real tensors");
    static constexpr int W = 1 + elem_t::intB + elem_t::fracB;
    using raw_t = typename detail::batch_word<W, Words>::type;
    static constexpr std::size_t per = TensorT::elemSize;

    explicit Qbatch(std::size_t count, Qmem mem = Qmem::pinned, int device = -1) : count_(count), mem_(mem), device_(device)
    {
        const std::size_t bytes = count * per * sizeof(raw_t);
        if (mem == Qmem::device) {
            detail::check(qb200_device_alloc(&p_, bytes, device));
        } else {
            if (mem != Qmem::pageable && qb200_host_alloc(&p_, bytes, mem == Qmem::write_combined ? QB_ALLOC_WRITE_COMBINED : QB_ALLOC_DEFAULT) != 0) {
                p_ = nullptr; mem_ = Qmem::pageable;                            // no GPU / no lockable memory
            }
            if (mem_ == Qmem::pageable) { p_ = ::operator new(bytes ? bytes : 16, std::align_val_t(64)); }
        }
    }
    ~Qbatch()
    {
        if (!p_) return;
        if (mem_ == Qmem::device) qb200_device_free(p_, device_);
        else if (mem_ != Qmem::pageable) qb200_host_free(p_);
        else ::operator delete(p_, std::align_val_t(64));
    }
    Qbatch(const Qbatch &) = delete;
    Qbatch &operator=(const Qbatch &) = delete;
    Qbatch(Qbatch &&o) noexcept : count_(o.count_), mem_(o.mem_), device_(o.device_), p_(o.p_) { o.p_ = nullptr; }

    std::size_t count() const { return count_; }
    std::size_t words() const { return count_ * per; }
    std::size_t bytes() const { return count_ * per * sizeof(raw_t); }
    Qmem memory() const { return mem_; }
    int device() const { return device_; }
    bool on_device() const { return mem_ == Qmem::device; }
    raw_t *raw() { return (raw_t *)p_; }
    const raw_t *raw() const { return (const raw_t *)p_; }
    static constexpr qb_qdesc desc() { qb_qdesc d = detail::lower(elem_t::spec); d.word_bytes = (uint8_t)sizeof(raw_t); return d; }

    // host access (host-resident batches only)
    raw_t &at(std::size_t inst, std::size_t linear) { host_only(); return raw()[inst * per + linear]; }
    TensorT get(std::size_t inst) const
    {
        host_only();
        TensorT t;
        for (std::size_t i = 0; i < per; i++) t[i].data.data = (typename elem_t::raw_t)raw()[inst * per + i];
        return t;
    }
    void set(std::size_t inst, const TensorT &t)
    {
        host_only();
        for (std::size_t i = 0; i < per; i++) raw()[inst * per + i] = (raw_t)t[i].data.data;
    }
    // counter-based synthetic inputs over the element type's full raw range, identical on host and device
    // (mirrors fill(), QuBLAS.h:526-536): word index = first_instance * elemSize + linear index
    void synth(std::uint64_t seed, std::uint64_t stream_id, std::uint64_t first_instance = 0)
    {
        qb_exec ex{ on_device() ? QB_MEM_DEVICE : QB_MEM_HOST, device_, nullptr, 0, 0 };
        detail::check(qb200_synth_fill(p_, words(), sizeof(raw_t), -(1LL << (W - 1)), (1LL << (W - 1)) - 1, seed, stream_id, first_instance * per, &ex));
    }
    // copies between batches of the same tensor type and word policy, in any direction (host <-> device)
    void copy_from(const Qbatch &src)
    {
        if (src.count_ != count_) throw std::invalid_argument("Qbatch::copy_from: batch sizes differ");
        detail::check(qb200_memcpy(p_, src.p_, bytes(), on_device() ? device_ : src.device_));
    }
private:
    void host_only() const { if (mem_ == Qmem::device) throw std::runtime_error("Qbatch: element access on a device-resident batch (copy it to a host batch first)"); }
    std::size_t count_;
    Qmem mem_;
    int device_;
    void *p_ = nullptr;
};

namespace detail {
template <class G, class AT, class BT, class CT, Qwords WA, Qwords WB, Qwords WC>
const qb_gemul_plan &batched_plan()
{
    using DA = dims_of<AT>; using DB = dims_of<BT>; using DC = dims_of<CT>;
    constexpr std::size_t M = G::ta ? DA::cols : DA::rows, K = G::ta ? DA::rows : DA::cols;
    constexpr std::size_t Kb = G::tb ? DB::cols : DB::rows, N = G::tb ? DB::rows : DB::cols;
    static_assert(K == Kb && DC::rows == M && DC::cols == N, "QgemulBatched: shapes do not match");
    using EA = typename DA::elem; using EB = typename DB::elem;
    static constexpr Spec mul = merge(EA::spec, EB::spec, parse<typename G::mul>(), true);
    using Low = reduce_lowering<typename G::add, K>;
    static constexpr qb_gemul_plan plan = [] {
        qb_gemul_plan p{};
        p.a = Qbatch<AT, WA>::desc(); p.b = Qbatch<BT, WB>::desc(); p.mul = lower(mul); p.c = Qbatch<CT, WC>::desc();
        p.add = Low::levels(mul);
        p.M = (uint32_t)M; p.N = (uint32_t)N; p.K = (uint32_t)K;
        p.trans_a = G::ta; p.trans_b = G::tb; p.tree_form = QB_TREE_VARIADIC;
        p.stride_a = (int64_t)(M * K); p.stride_b = (int64_t)(K * N); p.stride_c = (int64_t)(M * N);
        return p;
    }();
    return plan;
}
template <class A, class B, class C> qb_exec batched_exec(const A &a, const B &b, const C &c, int device)
{
    if (c.count() != a.count() || c.count() != b.count()) throw std::invalid_argument("QgemulBatched: batch sizes differ");
    if (a.on_device() != c.on_device() || b.on_device() != c.on_device()) throw std::invalid_argument("QgemulBatched: host and device batches cannot be mixed");
    if (c.on_device()) {
        if (a.device() != c.device() || b.device() != c.device()) throw std::invalid_argument("QgemulBatched: batches live on different devices");
        if (device != -1 && device != c.device()) throw std::invalid_argument("QgemulBatched: the batches live on another device");
        return qb_exec{ QB_MEM_DEVICE, c.device(), nullptr, 0, 0 };
    }
    return qb_exec{ QB_MEM_HOST, device, nullptr, 0, 0 };
}
}  // namespace detail

// device: CUDA ordinal, -1 = the current device, QB_DEVICE_ALL = shard the batch over every visible GPU (host batches)
template <class... Tags, class CT, class AT, class BT, Qwords WC, Qwords WA, Qwords WB>
void QgemulBatched(Qbatch<CT, WC> &C, const Qbatch<AT, WA> &A, const Qbatch<BT, WB> &B, int device = -1)
{
    const qb_gemul_plan &plan = detail::batched_plan<detail::gemul_tags<Tags...>, AT, BT, CT, WA, WB, WC>();
    qb_exec ex = detail::batched_exec(A, B, C, device);
    detail::check(qb200_gemul_batched(&plan, A.raw(), B.raw(), C.raw(), C.count(), &ex));
}

// Accuracy of a batched Qgemul result against the float64 GEMM of the same (dequantized) operands: SQNR sums and the
// maximum absolute error, reduced on the GPU in a fixed order.  With device-resident batches nothing but the 32-byte
// result crosses PCIe (Qu -> double, QuBLAS.h:2413-2416, on either side of the path).
struct Qerror {
    double sum_sq_err = 0, sum_sq_ref = 0, max_abs_err = 0;
    std::uint64_t n = 0;
    double sqnr_db() const { return 10.0 * std::log10(sum_sq_ref / sum_sq_err); }
};
template <class... Tags, class CT, class AT, class BT, Qwords WC, Qwords WA, Qwords WB>
Qerror QgemulError(const Qbatch<CT, WC> &C, const Qbatch<AT, WA> &A, const Qbatch<BT, WB> &B, int device = -1)
{
    const qb_gemul_plan &plan = detail::batched_plan<detail::gemul_tags<Tags...>, AT, BT, CT, WA, WB, WC>();
    qb_exec ex = detail::batched_exec(A, B, C, device);
    qb_err_stats st{};
    // one pass: the float64 reference of each instance lives in registers and is compared on the spot (host batches are staged once)
    detail::check(qb200_gemul_error_batched(&plan, A.raw(), B.raw(), C.raw(), C.count(), &st, &ex));
    return Qerror{ st.sum_sq_err, st.sum_sq_ref, st.max_abs_err, st.n };
}

}  // namespace QuBLAS
