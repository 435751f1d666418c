// short-circuit evaluation with side effects, a loop variable initialised from the outer one of the same name, nested
// inout calls, struct / array copies, float -> int truncation, int scalar times vector, ++ / -- on vectors, row-vector times
// matrix, continue inside do-while, signed integer hashing that overflows, swizzle of a call result, .length() of a member array
int calls = 0;
bool yes() { calls += 1; return true; }
bool no() { calls += 10; return false; }
int ihash(ivec2 p) { int n = p.x*1619 + p.y*31337; n = (n << 13) ^ n; return n*(n*n*15731 + 789221) + 1376312589; }
vec2 pair() { return vec2(0.25, 0.75); }
void twice(inout float x) { x *= 2.0; }
void nest(inout float x) { twice(x); twice(x); x += 1.0; }
struct Box { vec3 lo; float w[2]; };
void main() {
    bool r = no() && yes();          // yes() not called
    r = r || (yes() || no());         // no() not called
    int i = 3, total = 0;
    for (int i = i; i < 6; i++) total += i;          // inner i starts from the outer one
    float z = 1.5; nest(z);
    Box a = Box(vec3(1.0), float[2](2.0, 3.0)), b = a;
    b.lo.x = 9.0; b.w[1] = 7.0;                       // copies: a is untouched
    ivec3 t = ivec3(-1.5, 2.7, -0.2);
    vec3 v = vec3(gluv, 0.5);
    vec3 w = 2*v + 1.0/(v + 2.0) - v/2 + (-v);
    w++; --w; w *= 1.5;
    mat2 m = mat2(1.0, 2.0, 3.0, 4.0), mm = m*m + m/2.0 - mat2(0.5);
    vec2 rv = gluv*m, cv = m*gluv;
    int k = 0; do { k++; if (k == 2) continue; total += k; } while (k < 4);
    float h = float(ihash(ivec2(stxy)) & 0x7fffffff)/2147483647.0;
    fragColor = vec4(w*0.1 + vec3(rv - cv, pair().y) + 0.01*vec3(t) + 0.001*vec3(mm[0], mm[1].y),
                     0.01*float(calls) + 0.001*float(total) + 0.01*z + 0.001*(a.lo.x + a.w[1] + b.lo.x + b.w[1]) + 0.1*h + float(r)*0.1 + 0.01*float(a.w.length()));
}
