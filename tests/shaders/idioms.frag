// idioms of hand-written shaders (ShaderToy style): mainImage, rotation by `p.xy *= mat2`, several declarators per
// statement, comma steps, `type[n] name` arrays, arrays of structs, inout structs, while(true), hex literals, matrix
// element access, vector equality, nested ?:, early returns out of loops
#define ROT(a) mat2(cos(a), sin(a), -sin(a), cos(a))
#define SAT(x) clamp(x, 0., 1.)
#define ITER 6

struct Blob { vec2 center; float radius; vec3 tint; };
struct State { float energy; int hits; };

const float[4] WEIGHTS = float[4](.5, .25, .125, .0625);
const vec2 OFFSETS[3] = vec2[](vec2(-.6, .2), vec2(.5, -.3), vec2(.1, .6));

float field(vec2 p, Blob b) { return b.radius/(1e-3 + length(p - b.center)); }

void bump(inout State s, float amount) { s.energy += amount; s.hits += (amount > .8) ? 1 : 0; }

int firstAbove(float[4] values, float limit) {
    for (int i = 0; i < 4; i++) if (values[i] > limit) return i;
    return -1;
}

vec3 shade(vec2 p, float t) {
    Blob blobs[3];
    for (int i = 0; i < 3; ++i) blobs[i] = Blob(OFFSETS[i]*(1. + .1*sin(t + float(i))), .12 + .04*float(i), vec3(i == 0, i == 1, i == 2));
    State s = State(0., 0);
    vec3 col = vec3(0), sum = vec3(0.0), unused;
    for (int i = 0, j = 2; i < 3; i++, j--) {
        float f = field(p, blobs[i]), g = field(p.yx, blobs[j]);
        bump(s, f);
        col += blobs[i].tint*f + .1*g;
    }
    float values[4];
    for (int k = 0; k < 4; k++) values[k] = col[k % 3]*WEIGHTS[k];
    int first = firstAbove(values, .2);
    mat3 m = mat3(1.);
    m[1][2] = .25; m[2] = vec3(.1, .2, 1.);
    col = m*col;
    int n = 0;
    while (true) {
        if (n >= ITER || s.energy < .05) break;
        s.energy *= .5;
        n++;
    }
    float mask = float((0xF0 >> (n & 3)) & 0x3)/3.;
    col = (first < 0) ? col*.5 : (first == 0) ? col.zyx : (first == 1) ? col.yzx : col;
    if (vec2(s.hits, n) == vec2(1, ITER)) col = 1. - col;
    return SAT(col*(.5 + .5*mask));
}

void mainImage(out vec4 color, in vec2 coord) {
    vec2 p = (2.*coord - iResolution.xy)/iResolution.y;
    p.xy *= ROT(.3*iTime);
    p *= mat2(1.1, 0., 0., .9);
    vec3 c = shade(p, iTime);
    mat2 r = ROT(.5), ri = inverse(r);
    vec2 q = ri*(r*p) - p;
    c += 100.*abs(vec3(q, determinant(transpose(r)) - 1.));
    color = vec4(pow(c*.9 + .05, vec3(1./2.2)), 1);
}

void main() {
    mainImage(fragColor, fragCoord);
}
