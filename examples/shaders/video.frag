// The video texture over the whole frame, aspect-correct (written for this repository; compiled at run time by the
// GLSL -> CUDA translator: there is no ahead-of-time kernel for it).
void main() {
    GetCamera(iCamera);
    fragColor = vec4(stexture(iVideo, iCamera.stuv).rgb, 1.0);
}
