// Shadertoy idioms: function-like macros, const arrays, structs returned by value, do-while with break, .length(),
// uint arithmetic, vector ternaries, mainImage(out, in) called from main()
#define PI 3.14159265
#define rot(a) mat2(cos(a), -sin(a), sin(a), cos(a))
const int STEPS = 24;
const vec3 PALETTE[4] = vec3[4](vec3(0.5), vec3(0.5), vec3(1.0), vec3(0.0, 0.33, 0.67));
struct Hit { float d; int id; vec3 n; };

float hash21(vec2 p) { p = fract(p*vec2(123.34, 456.21)); p += dot(p, p + 45.32); return fract(p.x*p.y); }
vec3 pal(float t) { return PALETTE[0] + PALETTE[1]*cos(2.0*PI*(PALETTE[2]*t + PALETTE[3])); }
float noise(vec2 p) {
    vec2 i = floor(p), f = fract(p);
    f = f*f*(3.0 - 2.0*f);
    float a = hash21(i), b = hash21(i + vec2(1, 0)), c = hash21(i + vec2(0, 1)), d = hash21(i + vec2(1, 1));
    return mix(mix(a, b, f.x), mix(c, d, f.x), f.y);
}
float fbm(vec2 p) {
    float s = 0.0, a = 0.5;
    for (int i = 0; i < 5; i++) { s += a*noise(p); p = rot(0.5)*p*2.0 + 0.1; a *= 0.5; }
    return s;
}
Hit scene(vec3 p) {
    Hit h; h.d = 1e9; h.id = -1; h.n = vec3(0);
    float ds[3];
    ds[0] = length(p - vec3(0, 0, 3)) - 1.0;
    ds[1] = p.y + 1.0 + 0.1*sin(p.x*3.0);
    ds[2] = length(max(abs(p - vec3(1.5, 0, 3)) - vec3(0.4), 0.0)) - 0.05;
    for (int k = 0; k < ds.length(); ++k) if (ds[k] < h.d) { h.d = ds[k]; h.id = k; }
    return h;
}
void mainImage(out vec4 O, in vec2 U) {
    vec2 uv = (U - 0.5*iResolution.xy)/iResolution.y;
    uv *= rot(0.2*iTime);
    vec3 ro = vec3(0, 0, -1), rd = normalize(vec3(uv, 1.2));
    float t = 0.0; Hit h;
    int i = 0;
    do {
        h = scene(ro + rd*t);
        if (h.d < 1e-3) break;
        t += h.d;
        if (t > 20.0) { h.id = -1; break; }
    } while (++i < STEPS);
    vec3 col = h.id < 0 ? vec3(0.1, 0.2, 0.3)*fbm(uv*4.0) : pal(float(h.id)*0.3 + t*0.05);
    col.rg += 0.05*vec2(hash21(U), -hash21(U.yx));
    uint bits = uint(U.x) ^ (uint(U.y) << 3u);
    col.b *= float(bits % 7u)/7.0 + 0.5;
    O = vec4(pow(clamp(col, 0.0, 1.0), vec3(0.4545)), 1.0);
}
void main() { mainImage(fragColor, stxy); }
