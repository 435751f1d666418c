// implicit int -> float conversions (initialisers, arguments, returns), overloads by type, swizzle stores and chained
// swizzles in all three name sets, component / column stores into vectors and matrices, matrix resizing constructors,
// vector / matrix / struct equality, struct-valued ternary, shadowing, two-counter for with continue, builtins on mixed shapes
struct Pair { vec2 a; int n; };
float f(float x) { return x*2.0; }
float f(vec2 x) { return x.x - x.y; }
float f(int x) { return float(x) + 0.5; }
vec3 f(vec3 x, float s) { return x*s; }
float half_of(float x) { return x/2; }
float one() { return 1; }
Pair choose(bool c, Pair x, Pair y) { return c ? x : y; }
void main() {
    float x = 1;
    vec3 v = vec3(1);
    v.zx = gluv*2.0;
    v.yz *= 2.0;
    v[1] += 0.25;
    vec4 c = vec4(gluv, astuv);
    vec3 t = vec3(c);
    vec2 ch = c.xyz.zy + c.rgba.ab + c.stpq.ts;
    mat3 m3 = mat3(1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 10.0);
    mat2 m2 = mat2(m3);
    mat3 back = mat3(m2);
    m2[0][1] += gluv.x;
    m2[1] = vec2(0.5, -0.5);
    mat4 m4 = mat4(m2);
    bool same = (v == v) && (m2 != mat2(1.0)) && !(t == vec3(0));
    int k = int(astuv.x*3.9);
    uint wrap = uint(-1);
    ivec2 tr = ivec2(gluv*2.7);
    float conv = float(true) + float(int(true)) + float(bool(k)) + float(wrap > 5u) + float(tr.x) + 0.1*float(tr.y);
    Pair p = choose(gluv.x > 0.0, Pair(vec2(1, 2), 3), Pair(gluv, k));
    float over = f(x) + f(gluv) + f(k) + f(v, 0.5).y + half_of(3.0) + one();
    {
        float x = 5.0;          // shadows the outer x
        over += x;
    }
    for (int i = 0, j = 10; i < j; i += 2, j -= 1) { if (i == 4) continue; over += 0.01*float(i*j); }
    vec3 fn = max(v, 0.1) + clamp(t, -0.5, 0.5) + pow(abs(t) + 0.1, vec3(0.5, 1.5, 2.0)) + sqrt(abs(v))
            + vec3(length(-2.5), distance(gluv, astuv), sign(0.0)) + vec3(roundEven(2.5), trunc(-1.7), ceil(-0.2));
    float odd = float(isnan(sqrt(-abs(gluv.x) - 1.0))) + float(isinf(1.0/0.0)) + float(min(k, 2)) + float(abs(ivec2(-3, 2)).x);
    fragColor = vec4(fn*0.1 + 0.01*vec3(ch, conv) + 0.01*vec3(back[2][2], m4[3][3], m4[1][0]) + float(same)*0.1,
                     0.01*over + 0.1*odd + 0.01*float(p.n) + p.a.x*0.1);
}
