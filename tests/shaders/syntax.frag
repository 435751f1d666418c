// preprocessor (#ifdef / #elif defined() / #if ==), function-like macros, precision and parameter qualifiers, prototypes,
// arrays as values (returned, copied, struct members), comma in for, float loop counters, while, discard inside a helper,
// multiple declarators, unary / postfix operators, nested ternaries, v *= m against m*v, gl_FragCoord
precision highp float;
#define SQ(x) ((x)*(x))
#define LERP(a, b, t) mix(a, b, clamp(t, 0., 1.))
#define TWO_PI (2.*3.14159265)
#ifdef UNDEFINED_THING
#error should not be here
#elif defined(SQ) && !defined(NOPE)
#define MODE 2
#else
#define MODE 1
#endif
#if MODE == 2
const float GAIN = 1./3.;
#else
const float GAIN = 5.;
#endif

float ring(const in vec2 p, highp float r);          // prototype
float[3] weights(float t);                            // returns an array
struct Wave { float amp[3]; vec2 dir; };

float accumulate(Wave w, vec2 p) {
    float s = 0.;
    for (int i = 0, j = 2; i < 3; i++, j--) s += w.amp[i]*sin(dot(w.dir, p)*float(j + 1));
    return s;
}
void maybeDiscard(vec2 p) { if (SQ(p.x - 0.9) + SQ(p.y) < .002) discard; }
void main() {
    vec2 p = gluv, q, r = vec2(.5, -1e-1);
    q = -p.yx;
    maybeDiscard(p);
    float acc = 0., t;
    for (t = 0.; t < 1.; t += .25) acc += ring(p*(1. + t), .3 + .2*t);
    int n = 0;
    while (n < 4 && acc > float(n)*0.2) n++;
    float w[3] = weights(astuv.x);
    Wave wave = Wave(w, normalize(r));
    float a = accumulate(wave, q*TWO_PI);
    mat2 m = mat2(1., .5, -.5, 1.);
    p *= m;
    q = m*q;
    int k = n++ + 1;
    k += ~k & 3;
    bool flag = !(k > 2) || (n == 3 && p.x > 0.);
    vec3 col = flag ? (p.x > 0. ? vec3(1, .5, .2) : vec3(.2, .5, 1)) : vec3(.5);
    col = LERP(col, vec3(a*GAIN + .5), SQ(astuv.y));
    col += vec3(gl_FragCoord.xy/iResolution, float(k))*0.01;
    fragColor = vec4(col, acc*.25);
}
float ring(const in vec2 p, highp float r) { return smoothstep(.02, .0, abs(length(p) - r)); }
float[3] weights(float t) { return float[3](t, 1. - t, t*(1. - t)); }
