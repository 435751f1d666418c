// sfb200: scene=piano
//
// A minimal piano roll for ShaderPiano. The reference ships the module (shaderflow/piano/module.py) and its
// textures / uniforms but no fragment shader for it; this one is written for this repository, in the
// reference's GLSL dialect (it would run on the reference too), and transliterated to CUDA in
// shaderflow_b200/csrc/scenes.cuh (scene_piano).
//
//   iPianoRoll  (slot, pitch) -> (start, end, channel, velocity)    iPianoKeys (pitch, 0) -> key press 0..127
//   iPianoChan  (pitch, 0)    -> channel being played, -1 if none   iPianoDynamic -> visible pitch range

bool isBlackKey(int index) {
    int k = index - 12*(index/12);
    return (k == 1) || (k == 3) || (k == 6) || (k == 8) || (k == 10);
}

vec3 channelColor(float channel) {
    return hsv2rgb(0.6 + 0.9*channel, 0.75, 1.0);
}

void main() {
    vec3 rgb = vec3(0.06);
    fragColor = vec4(rgb, 1.0);

    // Horizontal axis: the dynamic note range plus extra keys on each side
    float lo = iPianoDynamic.x - iPianoExtra;
    float hi = iPianoDynamic.y + iPianoExtra + 1.0;
    float keyf = mix(lo, hi, astuv.x);
    float keyi = floor(keyf);
    float inkey = keyf - keyi;
    if ((keyi < 0.0) || (keyi > 127.0))
        return;
    int key = int(keyi);
    bool black = isBlackKey(key);

    if (astuv.y < iPianoHeight) {
        // The keyboard: pressed keys take the colour of the channel playing them
        float press = clamp(texelFetch(iPianoKeys, ivec2(key, 0), 0).r/128.0, 0.0, 1.0);
        float chan = texelFetch(iPianoChan, ivec2(key, 0), 0).r;
        vec3 base = black ? vec3(0.12) : vec3(0.92);
        if (black && (astuv.y < iPianoHeight*(1.0 - iPianoBlackRatio)))
            base = vec3(0.92);
        rgb = base;
        if (chan >= 0.0)
            rgb = mix(base, channelColor(chan), press);
        if (inkey < 0.06)
            rgb = rgb*0.55;
    } else {
        // The roll: height above the keyboard is time ahead of now
        float when = iTime + iPianoRollTime*(astuv.y - iPianoHeight)/(1.0 - iPianoHeight);
        rgb = black ? vec3(0.09) : vec3(0.12);
        for (int i = 0; i < iPianoLimit; i++) {
            vec4 note = texelFetch(iPianoRoll, ivec2(i, key), 0);
            if (note == vec4(0.0))
                break;
            if ((note.x <= when) && (when <= note.y)) {
                rgb = channelColor(note.z)*(0.35 + 0.65*note.w/128.0);
                if ((inkey < 0.08) || (inkey > 0.92))
                    rgb = rgb*0.5;
            }
        }
    }
    fragColor = vec4(rgb, 1.0);
}
