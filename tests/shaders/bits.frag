// integer and unsigned arithmetic, bit operations, ivec / uvec / bvec, while / do-while, switch with fall-through
uint pcg(uint v) {
    uint state = v*747796405u + 2891336453u;
    uint word = ((state >> ((state >> 28u) + 4u)) ^ state)*277803737u;
    return (word >> 22u) ^ word;
}

float unit(uint h) { return float(h & 0xffffu)/65535.0; }

int collatz(int n) {
    int steps = 0;
    while (n != 1 && steps < 40) {
        n = ((n & 1) == 0) ? n/2 : 3*n + 1;
        steps++;
    }
    return steps;
}

void main() {
    ivec2 cell = ivec2(floor(stxy/4.0));
    uvec2 u = uvec2(cell);
    uint h = pcg(u.x + pcg(u.y + uint(iFrame)));
    vec3 c = vec3(unit(h), unit(h >> 8u), unit(pcg(h)));
    int k = collatz(1 + cell.x % 27);
    bvec3 bright = greaterThan(c, vec3(0.5));
    float tone = 0.0;
    switch (k % 4) {
        case 0: tone = 0.2;
        case 1: tone += 0.1; break;
        case 2: tone = 0.7; break;
        default: tone = 1.0;
    }
    int i = 0;
    do { tone *= 0.97; i++; } while (i < cell.y % 5);
    if (any(bright) && !all(bright)) c = mix(c, vec3(tone), 0.5);
    if ((cell.x ^ cell.y) % 7 == 0) c = 1.0 - c;
    c.xz = c.zx;
    fragColor = vec4(c, float(k)/40.0);
}
