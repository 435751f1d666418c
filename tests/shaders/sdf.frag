// structs, out / inout parameters, matrices, const arrays, break / continue, nested calls
struct Ray { vec3 origin; vec3 direction; };
struct Hit { float distance; vec3 normal; int object; };

const vec3 CENTERS[3] = vec3[3](vec3(-1.2, 0.0, 4.0), vec3(0.9, 0.3, 3.5), vec3(0.0, -0.8, 5.0));
const float RADII[3] = float[](0.9, 0.6, 1.1);

mat3 look(float yaw, float pitch) {
    float cy = cos(yaw), sy = sin(yaw), cp = cos(pitch), sp = sin(pitch);
    mat3 ry = mat3(cy, 0, -sy, 0, 1, 0, sy, 0, cy);
    mat3 rx = mat3(1, 0, 0, 0, cp, sp, 0, -sp, cp);
    return ry*rx;
}

float scene(vec3 p, out int object) {
    float best = 1e9;
    object = -1;
    for (int i = 0; i < CENTERS.length(); i++) {
        float d = length(p - CENTERS[i]) - RADII[i];
        if (d >= best) continue;
        best = d;
        object = i;
    }
    float ground = p.y + 1.5;
    if (ground < best) { best = ground; object = 3; }
    return best;
}

bool march(Ray ray, inout Hit hit) {
    float t = 0.0;
    for (int i = 0; i < 48; i++) {
        int object;
        float d = scene(ray.origin + ray.direction*t, object);
        if (d < 0.002) {
            hit.distance = t;
            hit.object = object;
            vec2 e = vec2(0.01, 0);
            int unused;
            vec3 p = ray.origin + ray.direction*t;
            hit.normal = normalize(vec3(scene(p + e.xyy, unused) - scene(p - e.xyy, unused),
                                        scene(p + e.yxy, unused) - scene(p - e.yxy, unused),
                                        scene(p + e.yyx, unused) - scene(p - e.yyx, unused)));
            return true;
        }
        t += d;
        if (t > 20.0) break;
    }
    return false;
}

void main() {
    mat3 view = look(0.3*sin(iTime), 0.1);
    Ray ray = Ray(vec3(0, 0, 0), normalize(view*vec3(gluv, 1.5)));
    Hit hit = Hit(0.0, vec3(0), -1);
    vec3 color = vec3(0.05, 0.07, 0.1) + 0.1*ray.direction.y;
    if (march(ray, hit)) {
        vec3 light = normalize(vec3(0.5, 0.8, -0.4));
        float diffuse = max(dot(hit.normal, light), 0.0);
        vec3 r = reflect(ray.direction, hit.normal);
        float specular = pow(max(dot(r, light), 0.0), 16.0);
        vec3 albedo = (hit.object == 3) ? vec3(0.4) : vec3(0.9, 0.3, 0.2)*float(hit.object + 1)/3.0;
        color = albedo*(0.15 + diffuse) + specular;
        color *= 1.0/(1.0 + 0.02*hit.distance*hit.distance);
    }
    fragColor = vec4(color, 1);
}
