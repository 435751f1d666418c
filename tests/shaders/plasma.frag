// float loops, swizzles on both sides, vector builtins, a user uniform, ?: and compound assignment
uniform vec3 iTint;
#define LAYERS 5
#define WAVE(p, k) sin((p).x*(k) + iTime)*cos((p).y*(k) - iTime)

vec3 layer(vec2 p, float k) {
    vec3 c = vec3(WAVE(p, k), WAVE(p.yx, k*1.5), WAVE(p*0.5, k + 1.0));
    c.rg = c.gr*0.5 + 0.5;
    c.b *= 0.25;
    return c;
}

void main() {
    vec2 p = gluv*2.0;
    vec3 acc = vec3(0);
    float weight = 0.0;
    for (float k = 1.0; k < float(LAYERS) + 0.5; k += 1.0) {
        float w = 1.0/k;
        acc += layer(p, k)*w;
        weight += w;
        p = p.yx*vec2(1.1, -0.9) + 0.1;
    }
    acc /= weight;
    acc = mix(acc, iTint, smoothstep(0.6, 1.4, length(agluv)));
    float stripe = (mod(floor(stxy.x/8) + floor(stxy.y/8), 2) < 0.5) ? 1.0 : 0.92;
    fragColor = vec4(clamp(acc*stripe, 0, 1), step(0.25, astuv.x));
    fragColor.a = max(fragColor.a, 0.5);
}
