// glsl_rt.cuh — run-time library of the GLSL → CUDA translator (shaderflow_b200/glsl). Compiled by NVRTC only, as an
// in-memory header in front of the code the translator emits for a fragment shader this backend has no ahead-of-time
// kernel for (reference: ShaderProgram.compile hands ANY fragment text to the GL driver, shaderflow/shader.py:313-349).
//   g::vec<T,N>, g::mat<C,R>   GLSL's value types: component names by union, [] access, swizzles as swz<...>() /
//                              swz_set<...>(), constructors that flatten their arguments, implicit int → float
//   operators and builtins     GLSL 3.30 §5 / §8, generic over scalars and vectors; float32, one rounding per
//                              operation (the program is compiled with --fmad=false), sums left to right, min / max /
//                              clamp drop NaNs (fminf / fmaxf) — the rules oracle/glsl_exec.py states for the checker
//   texture access             the exact sampling path of sampler.cuh (point fetch + float32 bilinear weights)
// Self-contained: NVRTC has no host headers. Included after sfb200.h and render_params.h.
#pragma once

#define G_DEV __device__ __forceinline__

namespace g {

// ---------------------------------------------------------------------------------------------------------------------
// traits

template <bool B, class T = void> struct enable_if {};
template <class T> struct enable_if<true, T> { using type = T; };
template <class T> struct is_scalar { static constexpr bool value = false; };
template <> struct is_scalar<float> { static constexpr bool value = true; };
template <> struct is_scalar<double> { static constexpr bool value = true; };
template <> struct is_scalar<int> { static constexpr bool value = true; };
template <> struct is_scalar<unsigned> { static constexpr bool value = true; };
template <> struct is_scalar<bool> { static constexpr bool value = true; };
#define G_IF_SCALAR(S) class = typename enable_if<is_scalar<S>::value>::type

template <class T, int N> struct vec;
template <int C, int R> struct mat;

template <class X> struct info { using base = X; static constexpr int n = 1; };
template <class T, int N> struct info<vec<T, N>> { using base = T; static constexpr int n = N; };
template <class T, int N> struct make { using type = vec<T, N>; };
template <class T> struct make<T, 1> { using type = T; };
template <class A, class B> struct prom { using type = decltype(A() + B()); };
template <> struct prom<bool, bool> { using type = bool; };
template <class A> struct prom<A, double> { using type = float; };      // a stray double literal computes in float
template <class B> struct prom<double, B> { using type = float; };
template <> struct prom<double, double> { using type = float; };
constexpr int imax(int a, int b) { return a > b ? a : b; }

template <class S, G_IF_SCALAR(S)> G_DEV S get(const S& s, int) { return s; }
template <class T, int N> G_DEV T get(const vec<T, N>& v, int i) { return v.v[i]; }

// ---------------------------------------------------------------------------------------------------------------------
// vectors

#define G_VEC_COMMON(N) \
    G_DEV vec() { for (int i = 0; i < N; i++) v[i] = T(0); } \
    template <class U> G_DEV vec(const vec<U, N>& o) { for (int i = 0; i < N; i++) v[i] = T(o.v[i]); } \
    template <class... A> G_DEV explicit vec(const A&... a) { \
        if constexpr (sizeof...(A) == 1 && (is_scalar<A>::value && ...)) { for (int i = 0; i < N; i++) v[i] = T((a, ...)); } \
        else { int k = 0; (put(k, a), ...); } } \
    template <class S, G_IF_SCALAR(S)> G_DEV void put(int& k, const S& s) { if (k < N) v[k++] = T(s); } \
    template <class U, int M> G_DEV void put(int& k, const vec<U, M>& o) { for (int i = 0; i < M; i++) if (k < N) v[k++] = T(o.v[i]); } \
    template <int C, int R> G_DEV void put(int& k, const mat<C, R>& m) { for (int j = 0; j < C; j++) for (int i = 0; i < R; i++) if (k < N) v[k++] = T(m.c[j].v[i]); } \
    G_DEV T& operator[](int i) { return v[i]; } \
    G_DEV const T& operator[](int i) const { return v[i]; } \
    template <class X> G_DEV vec& operator+=(const X& o) { *this = vec(*this + o); return *this; } \
    template <class X> G_DEV vec& operator-=(const X& o) { *this = vec(*this - o); return *this; } \
    template <class X> G_DEV vec& operator*=(const X& o) { *this = vec(*this * o); return *this; } \
    template <class X> G_DEV vec& operator/=(const X& o) { *this = vec(*this / o); return *this; } \
    template <class X> G_DEV vec& operator%=(const X& o) { *this = vec(*this % o); return *this; } \
    template <class X> G_DEV vec& operator&=(const X& o) { *this = vec(*this & o); return *this; } \
    template <class X> G_DEV vec& operator|=(const X& o) { *this = vec(*this | o); return *this; } \
    template <class X> G_DEV vec& operator^=(const X& o) { *this = vec(*this ^ o); return *this; } \
    template <class X> G_DEV vec& operator<<=(const X& o) { *this = vec(*this << o); return *this; } \
    template <class X> G_DEV vec& operator>>=(const X& o) { *this = vec(*this >> o); return *this; } \
    G_DEV vec& operator++() { for (int i = 0; i < N; i++) v[i] = v[i] + T(1); return *this; } \
    G_DEV vec& operator--() { for (int i = 0; i < N; i++) v[i] = v[i] - T(1); return *this; } \
    G_DEV vec operator++(int) { vec o = *this; ++*this; return o; } \
    G_DEV vec operator--(int) { vec o = *this; --*this; return o; } \
    G_DEV int length() const { return N; }

template <class T> struct vec<T, 2> {
    union { struct { T x, y; }; struct { T r, g; }; struct { T s, t; }; T v[2]; };
    G_VEC_COMMON(2)
};
template <class T> struct vec<T, 3> {
    union { struct { T x, y, z; }; struct { T r, g, b; }; struct { T s, t, p; }; T v[3]; };
    G_VEC_COMMON(3)
};
template <class T> struct vec<T, 4> {
    union { struct { T x, y, z, w; }; struct { T r, g, b, a; }; struct { T s, t, p, q; }; T v[4]; };
    G_VEC_COMMON(4)
};

using vec2 = vec<float, 2>;   using vec3 = vec<float, 3>;   using vec4 = vec<float, 4>;
using ivec2 = vec<int, 2>;    using ivec3 = vec<int, 3>;    using ivec4 = vec<int, 4>;
using uvec2 = vec<unsigned, 2>; using uvec3 = vec<unsigned, 3>; using uvec4 = vec<unsigned, 4>;
using bvec2 = vec<bool, 2>;   using bvec3 = vec<bool, 3>;   using bvec4 = vec<bool, 4>;
using uint = unsigned;

#define G_BINOP(op) \
    template <class T, class U, int N> G_DEV auto operator op(const vec<T, N>& a, const vec<U, N>& b) { \
        using P = typename prom<T, U>::type; vec<P, N> r; for (int i = 0; i < N; i++) r.v[i] = P(a.v[i]) op P(b.v[i]); return r; } \
    template <class T, class S, int N, G_IF_SCALAR(S)> G_DEV auto operator op(const vec<T, N>& a, S b) { \
        using P = typename prom<T, S>::type; vec<P, N> r; for (int i = 0; i < N; i++) r.v[i] = P(a.v[i]) op P(b); return r; } \
    template <class S, class T, int N, G_IF_SCALAR(S)> G_DEV auto operator op(S a, const vec<T, N>& b) { \
        using P = typename prom<S, T>::type; vec<P, N> r; for (int i = 0; i < N; i++) r.v[i] = P(a) op P(b.v[i]); return r; }
G_BINOP(+) G_BINOP(-) G_BINOP(*) G_BINOP(/) G_BINOP(%) G_BINOP(&) G_BINOP(|) G_BINOP(^) G_BINOP(<<) G_BINOP(>>)
#undef G_BINOP

template <class T, int N> G_DEV vec<T, N> operator-(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i < N; i++) r.v[i] = -a.v[i]; return r; }
template <class T, int N> G_DEV vec<T, N> operator+(const vec<T, N>& a) { return a; }
template <class T, int N> G_DEV vec<T, N> operator~(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i < N; i++) r.v[i] = ~a.v[i]; return r; }
template <class T, class U, int N> G_DEV bool operator==(const vec<T, N>& a, const vec<U, N>& b) {
    bool e = true; for (int i = 0; i < N; i++) e = e && (a.v[i] == b.v[i]); return e; }
template <class T, class U, int N> G_DEV bool operator!=(const vec<T, N>& a, const vec<U, N>& b) { return !(a == b); }

// swizzles: v.xy → swz<0,1>(v); v.xy = e → swz_set<0,1>(v, e)
template <int... I, class V> G_DEV auto swz(const V& v) {
    using T = typename info<V>::base;
    if constexpr (sizeof...(I) == 1) return T((get(v, I), ...));
    else return vec<T, sizeof...(I)>(get(v, I)...);
}
template <int... I, class V, class X> G_DEV auto swz_set(V& v, const X& x) {
    using T = typename info<V>::base;
    int k = 0;
    ((v.v[I] = T(get(x, k++))), ...);
    return swz<I...>(v);
}

// ---------------------------------------------------------------------------------------------------------------------
// matrices: C columns of R rows (GLSL matCxR), column-major constructors

template <int C, int R> struct mat {
    vec<float, R> c[C];
    G_DEV mat() {}
    template <class... A> G_DEV explicit mat(const A&... a) {
        if constexpr (sizeof...(A) == 1 && (is_scalar<A>::value && ...)) {
            for (int j = 0; j < C; j++) for (int i = 0; i < R; i++) c[j].v[i] = (i == j) ? float((a, ...)) : 0.0f;
        } else { int k = 0; (put(k, a), ...); }
    }
    template <int C2, int R2> G_DEV explicit mat(const mat<C2, R2>& m) {
        for (int j = 0; j < C; j++) for (int i = 0; i < R; i++) c[j].v[i] = (j < C2 && i < R2) ? m.c[j].v[i] : ((i == j) ? 1.0f : 0.0f);
    }
    template <class S, G_IF_SCALAR(S)> G_DEV void put(int& k, const S& s) { if (k < C*R) { c[k/R].v[k % R] = float(s); k++; } }
    template <class U, int M> G_DEV void put(int& k, const vec<U, M>& o) { for (int i = 0; i < M; i++) put(k, o.v[i]); }
    G_DEV vec<float, R>& operator[](int j) { return c[j]; }
    G_DEV const vec<float, R>& operator[](int j) const { return c[j]; }
    template <class X> G_DEV mat& operator*=(const X& o) { *this = *this * o; return *this; }
    template <class X> G_DEV mat& operator+=(const X& o) { *this = *this + o; return *this; }
    template <class X> G_DEV mat& operator-=(const X& o) { *this = *this - o; return *this; }
    template <class X> G_DEV mat& operator/=(const X& o) { *this = *this / o; return *this; }
};
using mat2 = mat<2, 2>; using mat3 = mat<3, 3>; using mat4 = mat<4, 4>;
using mat2x2 = mat<2, 2>; using mat3x3 = mat<3, 3>; using mat4x4 = mat<4, 4>;
using mat2x3 = mat<2, 3>; using mat2x4 = mat<2, 4>; using mat3x2 = mat<3, 2>; using mat3x4 = mat<3, 4>;
using mat4x2 = mat<4, 2>; using mat4x3 = mat<4, 3>;

template <int C, int R> G_DEV vec<float, R> operator*(const mat<C, R>& m, const vec<float, C>& v) {
    vec<float, R> r;
    for (int i = 0; i < R; i++) { float s = m.c[0].v[i]*v.v[0]; for (int j = 1; j < C; j++) s = s + m.c[j].v[i]*v.v[j]; r.v[i] = s; }
    return r;
}
template <int C, int R> G_DEV vec<float, C> operator*(const vec<float, R>& v, const mat<C, R>& m) {
    vec<float, C> r;
    for (int j = 0; j < C; j++) { float s = v.v[0]*m.c[j].v[0]; for (int i = 1; i < R; i++) s = s + v.v[i]*m.c[j].v[i]; r.v[j] = s; }
    return r;
}
template <int K, int R, int C> G_DEV mat<C, R> operator*(const mat<K, R>& a, const mat<C, K>& b) {
    mat<C, R> r;
    for (int j = 0; j < C; j++) r.c[j] = a*b.c[j];
    return r;
}
#define G_MATOP(op) \
    template <int C, int R> G_DEV mat<C, R> operator op(const mat<C, R>& a, const mat<C, R>& b) { mat<C, R> r; for (int j = 0; j < C; j++) r.c[j] = a.c[j] op b.c[j]; return r; } \
    template <int C, int R, class S, G_IF_SCALAR(S)> G_DEV mat<C, R> operator op(const mat<C, R>& a, S b) { mat<C, R> r; for (int j = 0; j < C; j++) r.c[j] = a.c[j] op float(b); return r; } \
    template <int C, int R, class S, G_IF_SCALAR(S)> G_DEV mat<C, R> operator op(S a, const mat<C, R>& b) { mat<C, R> r; for (int j = 0; j < C; j++) r.c[j] = float(a) op b.c[j]; return r; }
G_MATOP(+) G_MATOP(-) G_MATOP(/)
#undef G_MATOP
template <int C, int R, class S, G_IF_SCALAR(S)> G_DEV mat<C, R> operator*(const mat<C, R>& a, S b) { mat<C, R> r; for (int j = 0; j < C; j++) r.c[j] = a.c[j]*float(b); return r; }
template <int C, int R, class S, G_IF_SCALAR(S)> G_DEV mat<C, R> operator*(S a, const mat<C, R>& b) { mat<C, R> r; for (int j = 0; j < C; j++) r.c[j] = float(a)*b.c[j]; return r; }
template <int C, int R> G_DEV mat<C, R> operator-(const mat<C, R>& a) { mat<C, R> r; for (int j = 0; j < C; j++) r.c[j] = -a.c[j]; return r; }
template <int C, int R> G_DEV bool operator==(const mat<C, R>& a, const mat<C, R>& b) { bool e = true; for (int j = 0; j < C; j++) e = e && (a.c[j] == b.c[j]); return e; }
template <int C, int R> G_DEV bool operator!=(const mat<C, R>& a, const mat<C, R>& b) { return !(a == b); }

template <int C, int R> G_DEV mat<R, C> transpose(const mat<C, R>& m) { mat<R, C> r; for (int j = 0; j < C; j++) for (int i = 0; i < R; i++) r.c[i].v[j] = m.c[j].v[i]; return r; }
template <int C, int R> G_DEV mat<C, R> matrixCompMult(const mat<C, R>& a, const mat<C, R>& b) { mat<C, R> r; for (int j = 0; j < C; j++) r.c[j] = a.c[j]*b.c[j]; return r; }
template <int R, int C> G_DEV mat<C, R> outerProduct(const vec<float, R>& a, const vec<float, C>& b) { mat<C, R> r; for (int j = 0; j < C; j++) r.c[j] = a*b.v[j]; return r; }
G_DEV float determinant(const mat2& m) { return m.c[0].v[0]*m.c[1].v[1] - m.c[1].v[0]*m.c[0].v[1]; }
G_DEV float determinant(const mat3& m) {
    return m.c[0].v[0]*(m.c[1].v[1]*m.c[2].v[2] - m.c[2].v[1]*m.c[1].v[2])
         - m.c[1].v[0]*(m.c[0].v[1]*m.c[2].v[2] - m.c[2].v[1]*m.c[0].v[2])
         + m.c[2].v[0]*(m.c[0].v[1]*m.c[1].v[2] - m.c[1].v[1]*m.c[0].v[2]);
}
G_DEV mat2 inverse(const mat2& m) { float d = 1.0f/determinant(m); return mat2(m.c[1].v[1]*d, -m.c[0].v[1]*d, -m.c[1].v[0]*d, m.c[0].v[0]*d); }
G_DEV mat3 inverse(const mat3& m) {
    const float a = m.c[0].v[0], b = m.c[1].v[0], c = m.c[2].v[0], d = m.c[0].v[1], e = m.c[1].v[1], f = m.c[2].v[1];
    const float g_ = m.c[0].v[2], h = m.c[1].v[2], i = m.c[2].v[2];
    const float A = e*i - f*h, B = f*g_ - d*i, C_ = d*h - e*g_, inv = 1.0f/(a*A + b*B + c*C_);
    return mat3(A*inv, B*inv, C_*inv, (c*h - b*i)*inv, (a*i - c*g_)*inv, (b*g_ - a*h)*inv, (b*f - c*e)*inv, (c*d - a*f)*inv, (a*e - b*d)*inv);
}
G_DEV float determinant(const mat4& m) {
    float d = 0.0f;
    for (int j = 0; j < 4; j++) {
        mat3 s; int cc = 0;
        for (int k = 0; k < 4; k++) { if (k == j) continue; for (int i = 1; i < 4; i++) s.c[cc].v[i - 1] = m.c[k].v[i]; cc++; }
        float t = m.c[j].v[0]*determinant(s);
        d = (j & 1) ? d - t : d + t;
    }
    return d;
}
G_DEV mat4 inverse(const mat4& m) {
    mat4 r; const float inv = 1.0f/determinant(m);
    for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
        mat3 s; int cc = 0;
        for (int k = 0; k < 4; k++) { if (k == j) continue; int rr = 0; for (int l = 0; l < 4; l++) { if (l == i) continue; s.c[cc].v[rr++] = m.c[k].v[l]; } cc++; }
        const float cof = determinant(s)*inv;
        r.c[i].v[j] = ((i + j) & 1) ? -cof : cof;
    }
    return r;
}

// .length() of arrays, vectors and matrices
// GLSL arrays are values: assigned, returned, passed by copy (GLSL 3.30 §4.1.9)
template <class T, int N> struct arr {
    T v[N];
    G_DEV T& operator[](int i) { return v[i]; }
    G_DEV const T& operator[](int i) const { return v[i]; }
};
template <class T, int N> G_DEV int length_of(T (&)[N]) { return N; }
template <class T, int N> G_DEV int length_of(const arr<T, N>&) { return N; }
template <class T, int N> G_DEV int length_of(const vec<T, N>&) { return N; }
template <int C, int R> G_DEV int length_of(const mat<C, R>&) { return C; }

// ---------------------------------------------------------------------------------------------------------------------
// builtins (GLSL 3.30 §8), generic over scalars and vectors

template <class F, class A> G_DEV auto map1f(const A& a) {
    constexpr int n = info<A>::n;
    if constexpr (n == 1) return F()(float(a));
    else { vec<decltype(F()(0.0f)), n> r; for (int i = 0; i < n; i++) r.v[i] = F()(float(a.v[i])); return r; }
}
template <class F, class A, class B> G_DEV auto map2f(const A& a, const B& b) {
    constexpr int n = imax(info<A>::n, info<B>::n);
    if constexpr (n == 1) return F()(float(a), float(b));
    else { vec<decltype(F()(0.0f, 0.0f)), n> r; for (int i = 0; i < n; i++) r.v[i] = F()(float(get(a, i)), float(get(b, i))); return r; }
}
template <class F, class A, class B, class C> G_DEV auto map3f(const A& a, const B& b, const C& c) {
    constexpr int n = imax(info<A>::n, imax(info<B>::n, info<C>::n));
    if constexpr (n == 1) return F()(float(a), float(b), float(c));
    else { vec<float, n> r; for (int i = 0; i < n; i++) r.v[i] = F()(float(get(a, i)), float(get(b, i)), float(get(c, i))); return r; }
}
// type-preserving (int stays int): abs, sign, min, max, clamp
template <class F, class A, class B> G_DEV auto map2t(const A& a, const B& b) {
    constexpr int n = imax(info<A>::n, info<B>::n);
    using P = typename prom<typename info<A>::base, typename info<B>::base>::type;
    if constexpr (n == 1) return F()(P(a), P(b));
    else { vec<P, n> r; for (int i = 0; i < n; i++) r.v[i] = F()(P(get(a, i)), P(get(b, i))); return r; }
}

#define G_F1(name, expr) struct f_##name { G_DEV auto operator()(float x) const { return expr; } }; \
    template <class A> G_DEV auto name(const A& a) { return map1f<f_##name>(a); }
#define G_F2(name, expr) struct f_##name { G_DEV auto operator()(float x, float y) const { return expr; } }; \
    template <class A, class B> G_DEV auto name(const A& a, const B& b) { return map2f<f_##name>(a, b); }
#define G_F3(name, expr) struct f_##name { G_DEV float operator()(float x, float y, float z) const { return expr; } }; \
    template <class A, class B, class C> G_DEV auto name(const A& a, const B& b, const C& c) { return map3f<f_##name>(a, b, c); }

G_F1(radians, x*3.14159265358979323846f/180.0f)  G_F1(degrees, x*180.0f/3.14159265358979323846f)
G_F1(sin, ::sinf(x))   G_F1(cos, ::cosf(x))   G_F1(tan, ::tanf(x))   G_F1(asin, ::asinf(x))   G_F1(acos, ::acosf(x))
G_F1(sinh, ::sinhf(x)) G_F1(cosh, ::coshf(x)) G_F1(tanh, ::tanhf(x)) G_F1(asinh, ::asinhf(x)) G_F1(acosh, ::acoshf(x)) G_F1(atanh, ::atanhf(x))
G_F1(exp, ::expf(x))   G_F1(log, ::logf(x))   G_F1(exp2, ::exp2f(x)) G_F1(log2, ::log2f(x))
G_F1(sqrt, ::sqrtf(x)) G_F1(inversesqrt, 1.0f/::sqrtf(x))
G_F1(floor, ::floorf(x)) G_F1(ceil, ::ceilf(x)) G_F1(trunc, ::truncf(x)) G_F1(round, ::rintf(x)) G_F1(roundEven, ::rintf(x))
G_F1(fract, x - ::floorf(x))
G_F1(isnan, bool(x != x)) G_F1(isinf, bool(::fabsf(x) == __int_as_float(0x7f800000)))
G_F2(pow, ::powf(x, y))
// modf / frexp / ldexp (GLSL 3.30 §8.3, 4.00 §8.3): the second result leaves through an out parameter
G_DEV float modf(float x, float& whole) { whole = ::truncf(x); return x - whole; }
template <int N> G_DEV vec<float, N> modf(const vec<float, N>& x, vec<float, N>& whole) {
    vec<float, N> r; for (int i = 0; i < N; i++) r.v[i] = modf(x.v[i], whole.v[i]); return r; }
G_DEV float frexp(float x, int& exponent) { return ::frexpf(x, &exponent); }
template <int N> G_DEV vec<float, N> frexp(const vec<float, N>& x, vec<int, N>& exponent) {
    vec<float, N> r; for (int i = 0; i < N; i++) r.v[i] = frexp(x.v[i], exponent.v[i]); return r; }
G_DEV float ldexp(float x, int exponent) { return ::ldexpf(x, exponent); }
template <int N> G_DEV vec<float, N> ldexp(const vec<float, N>& x, const vec<int, N>& exponent) {
    vec<float, N> r; for (int i = 0; i < N; i++) r.v[i] = ldexp(x.v[i], exponent.v[i]); return r; }
G_F2(mod, x - y*::floorf(x/y))
G_F2(step, (y < x) ? 0.0f : 1.0f)
G_F3(smoothstep_, (::fminf(::fmaxf((z - x)/(y - x), 0.0f), 1.0f)))
template <class A, class B, class C> G_DEV auto smoothstep(const A& e0, const B& e1, const C& x) {
    auto t = smoothstep_(e0, e1, x);
    return t*t*(3.0f - 2.0f*t);
}
struct f_atan2 { G_DEV float operator()(float y, float x) const { return ::atan2f(y, x); } };
struct f_atan1 { G_DEV float operator()(float x) const { return ::atanf(x); } };
template <class A> G_DEV auto atan(const A& a) { return map1f<f_atan1>(a); }
template <class A, class B> G_DEV auto atan(const A& y, const B& x) { return map2f<f_atan2>(y, x); }

struct f_min { template <class T> G_DEV T operator()(T a, T b) const { return (b < a) ? b : a; } G_DEV float operator()(float a, float b) const { return ::fminf(a, b); } };
struct f_max { template <class T> G_DEV T operator()(T a, T b) const { return (a < b) ? b : a; } G_DEV float operator()(float a, float b) const { return ::fmaxf(a, b); } };
template <class A, class B> G_DEV auto min(const A& a, const B& b) { return map2t<f_min>(a, b); }
template <class A, class B> G_DEV auto max(const A& a, const B& b) { return map2t<f_max>(a, b); }
template <class A, class B, class C> G_DEV auto clamp(const A& x, const B& lo, const C& hi) { return min(max(x, lo), hi); }
G_DEV float abs(float a) { return ::fabsf(a); }
G_DEV int abs(int a) { return (a < 0) ? -a : a; }
G_DEV float sign(float a) { return (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f); }
G_DEV int sign(int a) { return (a > 0) - (a < 0); }
template <class T, int N> G_DEV vec<T, N> abs(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i < N; i++) r.v[i] = abs(a.v[i]); return r; }
template <class T, int N> G_DEV vec<T, N> sign(const vec<T, N>& a) { vec<T, N> r; for (int i = 0; i < N; i++) r.v[i] = sign(a.v[i]); return r; }
// mix: x·(1−a) + y·a; a boolean selector picks per component
template <class A, class B, class C> G_DEV auto mix(const A& x, const B& y, const C& a) {
    constexpr int n = imax(info<A>::n, imax(info<B>::n, info<C>::n));
    using S = typename info<C>::base;
    if constexpr (n == 1) {
        if constexpr (S(2) == S(1)) return a ? float(y) : float(x);          // bool
        else return float(x)*(1.0f - float(a)) + float(y)*float(a);
    } else {
        vec<float, n> r;
        for (int i = 0; i < n; i++) r.v[i] = mix(float(get(x, i)), float(get(y, i)), get(a, i));
        return r;
    }
}

template <class T, class U, int N> G_DEV float dot(const vec<T, N>& a, const vec<U, N>& b) {
    float s = float(a.v[0])*float(b.v[0]); for (int i = 1; i < N; i++) s = s + float(a.v[i])*float(b.v[i]); return s; }
template <class A, class B, G_IF_SCALAR(A), G_IF_SCALAR(B)> G_DEV float dot(A a, B b) { return float(a)*float(b); }
template <class T, int N> G_DEV float length(const vec<T, N>& a) { return ::sqrtf(dot(a, a)); }
template <class A, G_IF_SCALAR(A)> G_DEV float length(A a) { return ::fabsf(float(a)); }
template <class A, class B> G_DEV float distance(const A& a, const B& b) { return length(a - b); }
template <class T, int N> G_DEV vec<float, N> normalize(const vec<T, N>& a) { return a/length(a); }
template <class A, G_IF_SCALAR(A)> G_DEV float normalize(A a) { return float(a)/::fabsf(float(a)); }
G_DEV vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }
template <class V> G_DEV V reflect(const V& i, const V& n) { return V(i - 2.0f*dot(n, i)*n); }
template <class V> G_DEV V faceforward(const V& n, const V& i, const V& nref) { return (dot(nref, i) < 0.0f) ? n : V(-n); }
template <class V, class S> G_DEV V refract(const V& i, const V& n, S eta_) {
    const float eta = float(eta_), d = dot(n, i), k = 1.0f - eta*eta*(1.0f - d*d);
    if (k < 0.0f) return V(0.0f);
    return V(eta*i - (eta*d + ::sqrtf(k))*n);
}

#define G_CMP(name, op) template <class T, class U, int N> G_DEV vec<bool, N> name(const vec<T, N>& a, const vec<U, N>& b) { \
    vec<bool, N> r; for (int i = 0; i < N; i++) r.v[i] = a.v[i] op b.v[i]; return r; }
G_CMP(lessThan, <) G_CMP(lessThanEqual, <=) G_CMP(greaterThan, >) G_CMP(greaterThanEqual, >=) G_CMP(equal, ==) G_CMP(notEqual, !=)
#undef G_CMP
template <int N> G_DEV bool any(const vec<bool, N>& a) { bool r = false; for (int i = 0; i < N; i++) r = r || a.v[i]; return r; }
template <int N> G_DEV bool all(const vec<bool, N>& a) { bool r = true; for (int i = 0; i < N; i++) r = r && a.v[i]; return r; }
template <int N> G_DEV vec<bool, N> not_(const vec<bool, N>& a) { vec<bool, N> r; for (int i = 0; i < N; i++) r.v[i] = !a.v[i]; return r; }

struct f_f2i { G_DEV int operator()(float x) const { return __float_as_int(x); } };
struct f_f2u { G_DEV unsigned operator()(float x) const { return __float_as_uint(x); } };
template <class A> G_DEV auto floatBitsToInt(const A& a) { return map1f<f_f2i>(a); }
template <class A> G_DEV auto floatBitsToUint(const A& a) { return map1f<f_f2u>(a); }
G_DEV float intBitsToFloat(int a) { return __int_as_float(a); }
G_DEV float uintBitsToFloat(unsigned a) { return __uint_as_float(a); }
template <class T, int N> G_DEV vec<float, N> intBitsToFloat(const vec<T, N>& a) { vec<float, N> r; for (int i = 0; i < N; i++) r.v[i] = __int_as_float(int(a.v[i])); return r; }
template <class T, int N> G_DEV vec<float, N> uintBitsToFloat(const vec<T, N>& a) { vec<float, N> r; for (int i = 0; i < N; i++) r.v[i] = __uint_as_float(unsigned(a.v[i])); return r; }

// ---------------------------------------------------------------------------------------------------------------------
// texture access: OpenGL 3.3 sampling rules, the exact path of sampler.cuh

struct sampler2D { const DevSampler* s; };

G_DEV int wrap_index(int i, int n, int repeat) {
    if (repeat) { i %= n; return (i < 0) ? i + n : i; }
    return ::min(::max(i, 0), n - 1);
}
#ifdef SFB_HOST_SHIM            // tests/test_glsl_host.py compiles this header for the host to check the translator without a GPU
G_DEV float half_bits_to_float(unsigned short h) { return sfb_host_half_to_float(h); }
#else
G_DEV float half_bits_to_float(unsigned short h) { float f; asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h)); return f; }
#endif

G_DEV vec4 texel_at(const DevSampler& s, int ix, int iy) {
    ix = wrap_index(ix, s.w, s.rx);
    iy = wrap_index(iy, s.h, s.ry);
    const size_t idx = size_t(iy)*size_t(s.w) + size_t(ix);
    vec4 t(0.0f, 0.0f, 0.0f, 1.0f);
    if (s.dtype == SFB_DTYPE_U8) {
        if (s.padded == 4) { const uchar4 c = __ldg(static_cast<const uchar4*>(s.lin) + idx); t = vec4(c.x/255.0f, c.y/255.0f, c.z/255.0f, c.w/255.0f); }
        else if (s.padded == 2) { const uchar2 c = __ldg(static_cast<const uchar2*>(s.lin) + idx); t.x = c.x/255.0f; t.y = c.y/255.0f; }
        else t.x = __ldg(static_cast<const unsigned char*>(s.lin) + idx)/255.0f;
    } else if (s.dtype == SFB_DTYPE_F32) {
        if (s.padded == 4) { const float4 c = __ldg(static_cast<const float4*>(s.lin) + idx); t = vec4(c.x, c.y, c.z, c.w); }
        else if (s.padded == 2) { const float2 c = __ldg(static_cast<const float2*>(s.lin) + idx); t.x = c.x; t.y = c.y; }
        else t.x = __ldg(static_cast<const float*>(s.lin) + idx);
    } else {
        const unsigned short* p = static_cast<const unsigned short*>(s.lin) + idx*size_t(s.padded);
        for (int k = 0; k < s.padded; k++) t.v[k] = half_bits_to_float(__ldg(p + k));
    }
    if (s.comps == 3) t.w = 1.0f;
    return t;
}
G_DEV vec4 texture(sampler2D t, vec2 uv) {
    const DevSampler& s = *t.s;
    const float u = uv.x*float(s.w), v = uv.y*float(s.h);
    if (s.filter == SFB_FILTER_NEAREST) return texel_at(s, int(::floorf(u)), int(::floorf(v)));
    const float ub = u - 0.5f, vb = v - 0.5f, fx = ::floorf(ub), fy = ::floorf(vb), a = ub - fx, b = vb - fy;
    const int i0 = int(fx), j0 = int(fy);
    const vec4 t00 = texel_at(s, i0, j0), t10 = texel_at(s, i0 + 1, j0), t01 = texel_at(s, i0, j0 + 1), t11 = texel_at(s, i0 + 1, j0 + 1);
    const vec4 top = t00*(1.0f - a) + t10*a, bot = t01*(1.0f - a) + t11*a;
    return top*(1.0f - b) + bot*b;
}
template <class S> G_DEV vec4 texture(sampler2D t, vec2 uv, S) { return texture(t, uv); }           // bias: one mip level
template <class S> G_DEV vec4 textureLod(sampler2D t, vec2 uv, S) { return texture(t, uv); }
template <class S> G_DEV ivec2 textureSize(sampler2D t, S) { return ivec2(t.s->w, t.s->h); }
// a fetch outside the image is undefined in GLSL 3.30; zeros here — what robust buffer access (and the GL drivers at
// hand) return, and the rule the ahead-of-time kernels and the checker follow (scenes.cuh texel_fetch_r, Life's borders)
template <class S> G_DEV vec4 texelFetch(sampler2D t, ivec2 p, S) {
    const DevSampler& s = *t.s;
    if (p.x < 0 || p.y < 0 || p.x >= s.w || p.y >= s.h) return vec4(0.0f);
    return texel_at(s, p.x, p.y);
}
// the offset / projective / explicit-gradient forms (GLSL 3.30 §8.7). Textures here have one level, so a gradient only
// ever selects that level; an offset is a whole number of texels added to the texel coordinates before wrapping
G_DEV vec4 textureOffset(sampler2D t, vec2 uv, ivec2 o) {
    const DevSampler& s = *t.s;
    const float u = uv.x*float(s.w), v = uv.y*float(s.h);
    if (s.filter == SFB_FILTER_NEAREST) return texel_at(s, int(::floorf(u)) + o.x, int(::floorf(v)) + o.y);
    const float ub = u - 0.5f, vb = v - 0.5f, fx = ::floorf(ub), fy = ::floorf(vb), a = ub - fx, b = vb - fy;
    const int i0 = int(fx) + o.x, j0 = int(fy) + o.y;
    const vec4 t00 = texel_at(s, i0, j0), t10 = texel_at(s, i0 + 1, j0), t01 = texel_at(s, i0, j0 + 1), t11 = texel_at(s, i0 + 1, j0 + 1);
    const vec4 top = t00*(1.0f - a) + t10*a, bot = t01*(1.0f - a) + t11*a;
    return top*(1.0f - b) + bot*b;
}
template <class S> G_DEV vec4 textureOffset(sampler2D t, vec2 uv, ivec2 o, S) { return textureOffset(t, uv, o); }
template <class S> G_DEV vec4 texelFetchOffset(sampler2D t, ivec2 p, S lod, ivec2 o) { return texelFetch(t, ivec2(p.x + o.x, p.y + o.y), lod); }
G_DEV vec4 textureGrad(sampler2D t, vec2 uv, vec2, vec2) { return texture(t, uv); }
G_DEV vec4 textureProj(sampler2D t, vec3 p) { return texture(t, vec2(p.x/p.z, p.y/p.z)); }
G_DEV vec4 textureProj(sampler2D t, vec4 p) { return texture(t, vec2(p.x/p.w, p.y/p.w)); }
template <class P, class S> G_DEV vec4 textureProj(sampler2D t, const P& p, S) { return textureProj(t, p); }

}  // namespace g
