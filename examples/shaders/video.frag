// The video texture over the whole frame, aspect-correct. Written for this repository; there is no ahead-of-time
// kernel for it: the GLSL -> CUDA translator compiles it when the scene compiles.
void main() {
    GetCamera(view);
    vec3 rgb = stexture(iVideo, view.stuv).rgb;
    fragColor = vec4(rgb, 1.0);
}
