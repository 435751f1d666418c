// nested structs, arrays of structs and floats, const struct, inout int, switch fall-through with return, mat3 / mat4
// constructors and column access, inverse / transpose / determinant / outerProduct, bvec mix, any / all, reflect / refract,
// integer division, floatBitsToUint
struct Material { vec3 albedo; float rough; };
struct Light { vec3 dir; Material tint; };
const Material GOLD = Material(vec3(1.0, 0.77, 0.34), 0.3);
Light lights[2];
float table[4] = float[](0.1, 0.2, 0.4, 0.8);
int counter = 0;

mat3 lookAt(vec3 f) {
    f = normalize(f);
    vec3 r = normalize(cross(vec3(0, 1, 0), f)), u = cross(f, r);
    return mat3(r, u, f);
}
float sum(float xs[4]) { float s = 0.0; for (int i = 0; i < 4; i++) s += xs[i]; return s; }
void bump(inout int c, int by) { if (by == 0) return; c += by; }
vec3 shade(vec3 n, vec3 v, Light l) {
    vec3 h = normalize(l.dir + v);
    float ndl = max(dot(n, l.dir), 0.0), ndh = max(dot(n, h), 0.0);
    return l.tint.albedo*(ndl + pow(ndh, 2.0/(l.tint.rough*l.tint.rough + 1e-3)));
}
int classify(float x) {
    switch (int(floor(x*4.0))) {
        case 0: return 1;
        case 1:
        case 2: bump(counter, 2); return 2;
        default: break;
    }
    return 0;
}
void main() {
    lights[0] = Light(normalize(vec3(1, 1, -1)), GOLD);
    lights[1] = Light(normalize(vec3(-1, 0.5, -0.5)), Material(vec3(0.2, 0.3, 0.9), 0.6));
    mat3 cam = lookAt(vec3(0.2*sin(iTime), 0.1, 1.0));
    vec3 rd = cam*normalize(vec3(gluv, 1.5));
    mat3 inv = inverse(cam), tr = transpose(cam);
    float ortho = length(inv[0] - tr[0]) + abs(inv[1][2] - tr[1][2]) + length(inv[2].yz - tr[2].yz);
    vec3 n = normalize(vec3(gluv, -sqrt(max(0.0, 1.0 - dot(gluv, gluv)))));
    vec3 col = vec3(0);
    for (int i = 0; i < 2; i++) col += shade(n, -rd, lights[i]);
    col *= sum(table)/1.5;
    bvec3 dark = lessThan(col, vec3(0.2));
    col = mix(col, vec3(0.2) + 0.1*reflect(rd, n), dark);
    if (any(dark) && !all(dark)) col.r += 0.05;
    vec3 refr = refract(rd, n, 0.8);
    int kind = classify(astuv.x);
    col += 0.1*refr*float(kind) + 0.01*float(counter);
    int q = int(stxy.x) - 20, m = q/7 + (q - (q/7)*7);
    uint fb = floatBitsToUint(astuv.y) >> 20u;
    col.g += 0.001*float(m) + 0.0001*float(fb & 255u);
    mat2 o = outerProduct(gluv, vec2(0.5, -0.25));
    col.b += determinant(o) + o[1].x*0.1 + ortho;
    mat4 big = mat4(1.0); big[3] = vec4(gluv, 0.0, 1.0);
    vec4 moved = big*vec4(col, 1.0);
    fragColor = vec4(inversesqrt(1.0 + moved.xyz*moved.xyz)*moved.xyz, smoothstep(vec2(0.0), vec2(1.0), astuv).x);
}
