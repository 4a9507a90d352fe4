// probe_kernels.cl - OUR test wrappers, appended after the reference's translation unit at JIT time.
// They only *call* the reference's own device functions (closestIntersect, Camera_pinHole, ...), so the
// first-hit buffers come from the unmodified reference traversal/shading code.  Test infrastructure only.

__kernel void probe_first_hit(
    __global const int* projectorType, __global const float* cameraSettings,
    __global const int* octreeDepth, __global const int* octreeData,
    __global const int* bPalette, __global const int* quadModels, __global const int* aabbModels,
    __global const int* worldBvhData, __global const int* actorBvhData, __global const int* bvhTrigs,
    image2d_array_t textureAtlas, __global const int* matPalette,
    __global const int* randomSeed, __global const int* width, __global const int* height,
    __global int* outHit, __global int* outMaterial, __global float* outT,
    __global float* outNormal, __global float* outColor, __global float* outRay)
{
    int gid = get_global_id(0);
    Pixel pixel = Pixel_new(gid);
    Ray ray = Ray_new(&pixel);
    IntersectionRecord record = IntersectionRecord_new(&ray);
    MaterialPalette materialPalette = MaterialPalette_new(matPalette);
    Octree octree = Octree_create(octreeData, *octreeDepth);
    Bvh worldBvh = Bvh_new(worldBvhData, bvhTrigs, &materialPalette);
    Bvh actorBvh = Bvh_new(actorBvhData, bvhTrigs, &materialPalette);
    BlockPalette blockPalette = BlockPalette_new(bPalette, quadModels, aabbModels, &materialPalette);

    unsigned int rs = *randomSeed + gid;
    Random_nextState(&rs);

    if (*projectorType != -1) {
        // same expressions as the render kernel's camera block, so the JIT sees the same arithmetic
        float3 cameraPos = vload3(0, cameraSettings);
        float3 m1s = vload3(1, cameraSettings);
        float3 m2s = vload3(2, cameraSettings);
        float3 m3s = vload3(3, cameraSettings);
        float halfWidth = (*width) / (2.0 * (*height));
        float invHeight = 1.0 / (*height);
        float x = -halfWidth + ((pixel.index % (*width)) + Random_nextFloat(&rs)) * invHeight;
        float y = -0.5 + ((pixel.index / (*width)) + Random_nextFloat(&rs)) * invHeight;
        if (*projectorType == 0) Camera_pinHole(x, y, &rs, &ray.origin, &ray.direction, cameraSettings+12);
        ray.direction = (float3) (dot(m1s, ray.direction), dot(m2s, ray.direction), dot(m3s, ray.direction));
        ray.origin = (float3) (dot(m1s, ray.origin), dot(m2s, ray.origin), dot(m3s, ray.origin));
        ray.origin += cameraPos;
    } else {
        Camera_preGenerated(&ray, cameraSettings);
    }
    vstore3(ray.origin, gid * 2, outRay);
    vstore3(ray.direction, gid * 2 + 1, outRay);

    bool hit = closestIntersect(&record, &octree, &blockPalette, textureAtlas, 256, &worldBvh, &actorBvh);
    outHit[gid] = hit ? 1 : 0;
    outMaterial[gid] = hit ? record.material : 0;
    outT[gid] = hit ? record.distance : HUGE_VALF;
    vstore3(hit ? record.normal : (float3)(0, 0, 0), gid, outNormal);
    vstore4(hit ? record.color : (float4)(0, 0, 0, 0), gid, outColor);
}

// evaluates the runtime's builtins so their deviation from the oracle's deterministic math can be reported
__kernel void probe_math(__global const float* x, __global const float* y, __global float* out, int fn)
{
    int i = get_global_id(0);
    float r;
    switch (fn) {
        case 0: r = sin(x[i]); break;
        case 1: r = cos(x[i]); break;
        case 2: r = atan2(x[i], y[i]); break;
        case 3: r = asin(x[i]); break;
        case 4: r = acos(x[i]); break;
        case 5: r = 1 / x[i]; break;
        case 6: r = sqrt(x[i]); break;
        case 7: r = x[i] / y[i]; break;
        default: r = normalize((float3)(x[i], y[i], 1.0f)).x; break;
    }
    out[i] = r;
}

// vector builtins as the runtime expands them (dot / cross / normalize / length), for the arithmetic contract
__kernel void probe_vec(__global const float* a, __global const float* b, __global float* out)
{
    int i = get_global_id(0);
    float3 x = vload3(i, a);
    float3 y = vload3(i, b);
    float3 c = cross(x, y);
    float3 n = normalize(x);
    vstore3(c, i * 3, out);
    vstore3(n, i * 3 + 1, out);
    out[i * 9 + 6] = dot(x, y);
    out[i * 9 + 7] = length(x);
    out[i * 9 + 8] = 1 / sqrt(dot(x, x));
}
