// textureOffset / texelFetchOffset / textureProj / textureGrad on LINEAR-repeat and NEAREST-clamp samplers, and the
// builtins with an out parameter: modf, frexp (+ ldexp), also through a swizzle
uniform sampler2D picture;
uniform sampler2D table;

void main() {
    vec2 uv = astuv*1.25 - 0.125;
    // a 3 x 3 box from offsets: what a blur written with textureOffset does
    vec3 box = vec3(0.0);
    box += textureOffset(picture, uv, ivec2(-1, -1)).rgb + textureOffset(picture, uv, ivec2(0, -1)).rgb + textureOffset(picture, uv, ivec2(1, -1)).rgb;
    box += textureOffset(picture, uv, ivec2(-1,  0)).rgb + texture(picture, uv).rgb                     + textureOffset(picture, uv, ivec2(1,  0)).rgb;
    box += textureOffset(picture, uv, ivec2(-1,  1)).rgb + textureOffset(picture, uv, ivec2(0,  1)).rgb + textureOffset(picture, uv, ivec2(1,  1)).rgb;
    box /= 9.0;
    vec4 stepped = textureOffset(table, vec2(astuv.x, 0.5), ivec2(2, 0));                 // NEAREST, clamped at the edge
    vec4 fetched = texelFetchOffset(table, ivec2(int(astuv.y*4.0), 0), 0, ivec2(3, 0));           // stays inside: out of range is undefined
    vec3 projected = textureProj(picture, vec3(uv*(1.5 + astuv.x), 1.5 + astuv.x)).rgb
                   + textureProj(picture, vec4(uv*2.0, 7.0, 2.0)).rgb
                   - 2.0*textureGrad(picture, uv, vec2(0.01, 0.0), vec2(0.0, 0.01)).rgb;      // == 0 up to rounding

    vec2 whole;
    vec2 part = modf(gluv*3.7, whole);
    vec4 store = vec4(0.0);
    float f = modf(glxy.x*0.37, store.z);
    ivec2 e;
    vec2 m = frexp(vec2(glxy.x*0.01 + 2.0, gluv.y*300.0), e);
    int e1;
    float m1 = frexp(0.001 + abs(gluv.x), e1);
    float back = ldexp(m1, e1) - (0.001 + abs(gluv.x));                                       // exactly 0
    vec2 scaled = ldexp(m, ivec2(3, -2));

    fragColor = vec4(box + 0.05*stepped.rgb + 0.05*fetched.a + projected,
                     0.1*part.x + 0.01*whole.y + 0.1*f + 0.001*store.z + 0.01*float(e.x + e.y + e1) + m.x + back + 0.01*scaled.y);
}
