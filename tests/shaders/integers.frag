// integer hashing with uint overflow (pcg3d, wang), vector compound bit operators, int / ivec division, modulo and shifts of
// negative numbers, uint wrap-around, equal() on ivec, mix with bvec, mod / step / exp2 / log2 / atan on mixed shapes
uvec3 pcg3d(uvec3 v) {
    v = v*1664525u + 1013904223u;
    v.x += v.y*v.z; v.y += v.z*v.x; v.z += v.x*v.y;
    v ^= v >> 16u;
    v.x += v.y*v.z; v.y += v.z*v.x; v.z += v.x*v.y;
    return v;
}
vec3 hash33(vec3 p) { return vec3(pcg3d(uvec3(ivec3(floor(p)) + 1000)))*(1.0/float(0xffffffffu)); }
uint wang(uint s) { s = (s ^ 61u) ^ (s >> 16); s *= 9u; s = s ^ (s >> 4); s *= 0x27d4eb2du; s = s ^ (s >> 15); return s; }
int imod(int a, int b) { return a - b*(a/b); }
void main() {
    ivec2 cell = ivec2(floor(stxy/4.0));
    ivec2 wrapped = ivec2(imod(cell.x, 5), cell.y % 3);
    vec3 h = hash33(vec3(vec2(cell), float(iFrame)));
    uint w = wang(uint(cell.x) + 64u*uint(cell.y));
    float r = float(w & 0xffffu)/65535.0, g = float((w >> 16) & 0xffu)/255.0;
    ivec3 q = clamp(ivec3(h*10.0) - 5, ivec3(-3), ivec3(3));
    ivec3 a = abs(q)*sign(q) + min(q, ivec3(1)) - max(q, -2);
    bvec2 odd = equal(wrapped & 1, ivec2(1));
    vec2 sel = mix(vec2(0.25), vec2(0.75), odd);
    vec3 m = mod(vec3(gluv*3.0, iTime), 1.5) + mod(-gluv.xyx, vec3(0.7, 0.9, 1.1));
    float ang = atan(gluv.y, gluv.x) + atan(gluv.x*0.5);
    vec3 e = exp2(-abs(vec3(gluv, 0.5))) + log2(1.0 + abs(m)) + vec3(step(0.5, h.x), step(vec2(0.3, 0.6), h.yz));
    uvec2 packed = uvec2(cell) << uvec2(1u, 2u);
    int neg = -7/2 + (-7 % 3) + (7 >> 1) + (-8 >> 1) + int(3u - 5u > 10u);
    fragColor = vec4(h*0.3 + 0.2*vec3(r, g, sel.x) + 0.01*vec3(a) + 0.05*e + 0.02*m,
                     0.1*ang + 0.001*float(packed.x + packed.y) + 0.01*float(neg) + sel.y*0.1);
}
