// texture(), texelFetch(), textureSize() on LINEAR / NEAREST, repeat / clamp samplers; discard; global state
uniform sampler2D picture;
uniform sampler2D table;
uniform float iGain;
float calls = 0.0;

vec3 fetch(vec2 uv) {
    calls += 1.0;
    return texture(picture, uv).rgb;
}

void main() {
    if (length(agluv) > 1.35) discard;
    ivec2 size = textureSize(picture, 0);
    vec2 texel = 1.0/vec2(size);
    vec3 c = fetch(astuv*1.5 - 0.25);
    for (int dx = -1; dx <= 1; dx += 2)
        for (int dy = -1; dy <= 1; dy += 2)
            c += 0.25*fetch(astuv*1.5 - 0.25 + texel*vec2(dx, dy)*2.5);
    c /= 2.0;
    vec4 lut = texelFetch(table, ivec2(int(astuv.x*float(textureSize(table, 0).x)), 0), 0);
    c = mix(c, c*lut.rgb*iGain, 0.5);
    vec4 point = texture(table, vec2(astuv.y, 0.5));
    fragColor = vec4(c + 0.1*point.a, calls/5.0);
}
