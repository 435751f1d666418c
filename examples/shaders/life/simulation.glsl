// sfb200: scene=life_simulation
//
// Placeholder for the reference's examples/basic/shaders/life/simulation.glsl (not redistributed here).
// The CUDA backend renders this scene with the device function transliterated from that file
// (shaderflow_b200/csrc/scenes.cuh); the directive above selects it. Pointing `shader.fragment` at
// the reference's own file works too: it is recognised by content hash (shaderflow_b200/registry.py).
