package dev.thatredox.chunkynative.cuda;

import dev.thatredox.chunkynative.common.export.AbstractSceneLoader;
import dev.thatredox.chunkynative.common.export.ResourcePalette;
import dev.thatredox.chunkynative.common.export.models.PackedAabbModel;
import dev.thatredox.chunkynative.common.export.models.PackedQuadModel;
import dev.thatredox.chunkynative.common.export.models.PackedTriangleModel;
import dev.thatredox.chunkynative.common.export.primitives.PackedBlock;
import dev.thatredox.chunkynative.common.export.primitives.PackedMaterial;
import dev.thatredox.chunkynative.common.export.texture.AbstractTextureLoader;
import dev.thatredox.chunkynative.common.state.SkyState;
import dev.thatredox.chunkynative.util.Reflection;
import dev.thatredox.chunkynative.util.Util;
import se.llbit.chunky.renderer.ResetReason;
import se.llbit.chunky.renderer.scene.Camera;
import se.llbit.chunky.renderer.scene.Scene;
import se.llbit.chunky.renderer.scene.Sky;
import se.llbit.chunky.renderer.scene.SkyCache;
import org.apache.commons.math3.util.FastMath;
import se.llbit.math.Matrix3;
import se.llbit.math.Ray;
import se.llbit.math.Vector3;

import java.util.concurrent.ThreadLocalRandom;
import java.util.concurrent.locks.Lock;
import java.util.stream.IntStream;

/**
 * Replacement for opencl.renderer.ClSceneLoader (+ scene/ClSky, scene/ClCamera): the device-agnostic packers in
 * common/export are reused unchanged and still define every layout; the sinks change from ClIntBuffer / cl_mem / cl images
 * to ccu_scene_set_* calls, and the scene is committed once everything is in place (ccu_scene_commit builds the traversal
 * layouts).  Mirrors chunkyclplugin_b200/renderer.py::CudaSceneLoader / CudaCamera, the variant exercised by the tests.
 * NOT COMPILED in the build image (no JDK, no chunky-core jar).
 */
public class CudaSceneLoader extends AbstractSceneLoader {
    protected final ChunkyCu.Context ctx;
    protected SkyState skyState = null;
    private boolean skyLoaded = false;

    // produced by loadOctree(), consumed at the end of load()
    private int[] pendingOctree = null;
    private int pendingDepth = 0;
    private boolean palettesDirty = false;

    public CudaSceneLoader(ChunkyCu.Context ctx) { this.ctx = ctx; }
    public CudaSceneLoader() { this(CudaRendererInstance.get().context); }

    public ChunkyCu.Context context() { return ctx; }

    @Override
    public boolean ensureLoad(Scene scene) { return this.ensureLoad(scene, !skyLoaded); }        // ClSceneLoader.java:34-36

    @Override
    public boolean load(int modCount, ResetReason resetReason, Scene scene) {
        boolean skyChanged = false;
        if (this.modCount != modCount) {                                                       // ClSceneLoader.java:39-49
            SkyState now = new SkyState(scene.sky(), scene.sun());
            if (!now.equals(skyState)) { skyState = now; skyChanged = true; }
        }
        Object before = this.blockPalette;
        if (!super.load(modCount, resetReason, scene)) return false;
        palettesDirty |= this.blockPalette != before;

        if (!palettesDirty && pendingOctree == null && !skyChanged) return true;
        ctx.sceneBegin();
        if (palettesDirty) {
            // the texture atlas was written by CudaTextureLoader.buildTextures() during super.load()
            upload((CudaPackedResourcePalette<?>) this.blockPalette, ctx::setBlockPalette);
            upload((CudaPackedResourcePalette<?>) this.materialPalette.palette, ctx::setMaterialPalette);
            upload((CudaPackedResourcePalette<?>) this.aabbPalette, ctx::setAabbModels);
            upload((CudaPackedResourcePalette<?>) this.quadPalette, ctx::setQuadModels);
            upload((CudaPackedResourcePalette<?>) this.trigPalette, ctx::setTriangles);
            ctx.setWorldBvh(this.worldBvh);
            ctx.setActorBvh(this.actorBvh);
            ctx.setSun(this.packedSun.pack().toIntArray());                                     // ClSceneLoader.java:148-150
            palettesDirty = false;
        }
        if (skyChanged || !skyLoaded) {
            uploadSky(scene);
            skyLoaded = true;
        }
        if (pendingOctree != null) {
            ctx.setOctree(pendingOctree, pendingDepth);
            pendingOctree = null;
        }
        ctx.sceneCommit();
        return true;
    }

    private interface WordSink { void accept(int[] words, int n); }

    private static void upload(CudaPackedResourcePalette<?> palette, WordSink sink) { sink.accept(palette.elements(), palette.size()); }

    @Override
    protected boolean loadOctree(int[] octree, int depth, int[] blockMapping, ResourcePalette<PackedBlock> blockPalette) {
        // leaf remap of ClSceneLoader.java:56-58: -type -> -blockMapping[type]; types outside the palette (ANY_TYPE) stay as they are
        int[] mapped = new int[octree.length];
        for (int i = 0; i < octree.length; i++) {
            int w = octree[i];
            mapped[i] = (w > 0 || -w >= blockMapping.length) ? w : -blockMapping[-w];
        }
        pendingOctree = mapped;
        pendingDepth = depth;
        return true;
    }

    @Override protected AbstractTextureLoader createTextureLoader() { return new CudaTextureLoader(ctx); }
    @Override protected ResourcePalette<PackedBlock> createBlockPalette() { return new CudaPackedResourcePalette<>(); }
    @Override protected ResourcePalette<PackedMaterial> createMaterialPalette() { return new CudaPackedResourcePalette<>(); }
    @Override protected ResourcePalette<PackedAabbModel> createAabbModelPalette() { return new CudaPackedResourcePalette<>(); }
    @Override protected ResourcePalette<PackedQuadModel> createQuadModelPalette() { return new CudaPackedResourcePalette<>(); }
    @Override protected ResourcePalette<PackedTriangleModel> createTriangleModelPalette() { return new CudaPackedResourcePalette<>(); }

    /** ClSky.java:23-62: the sky baked into a res x res RGBA8 equirectangular table, plus the sun intensity. */
    private void uploadSky(Scene scene) {
        int res = skyResolution(scene);
        byte[] texels = new byte[res * res * 4];
        Ray ray = new Ray();
        // FastMath and the i-outer / j-inner order of ClSky.java:43-58: the texel bytes are truncated, so the same
        // transcendental implementation is needed to land on the same byte at a truncation edge
        for (int i = 0; i < res; i++) {
            for (int j = 0; j < res; j++) {
                double theta = ((double) i / res) * 2 * FastMath.PI;
                double phi = ((double) j / res) * FastMath.PI - FastMath.PI / 2;
                double r = FastMath.cos(phi);
                ray.d.set(FastMath.cos(theta) * r, FastMath.sin(phi), FastMath.sin(theta) * r);
                scene.sky().getSkyColor(ray, false);
                int at = 4 * (j * res + i);
                texels[at] = (byte) (ray.color.x * 255);
                texels[at + 1] = (byte) (ray.color.y * 255);
                texels[at + 2] = (byte) (ray.color.z * 255);
                texels[at + 3] = (byte) 255;
            }
        }
        ctx.setSky(texels, res, (float) scene.sun().getIntensity());
    }

    private static int skyResolution(Scene scene) {                                              // ClSky.java:64-76
        try {
            Sky sky = scene.sky();
            java.lang.reflect.Field f = sky.getClass().getDeclaredField("skyCache");
            f.setAccessible(true);
            return ((SkyCache) f.get(sky)).getSkyResolution();
        } catch (NoSuchFieldException | IllegalAccessException e) {
            return 128;
        }
    }

    /**
     * ClCamera.java:33-105: 15 floats for the pinhole projector (position - origin, row-major transform, aperture, subject
     * distance, fovTan), or 6 floats per pixel of rays generated by Chunky's own projector for every other projection mode.
     */
    public boolean uploadCamera(Scene scene, Lock renderLock, boolean jitter) {      // returns ClCamera.needGenerate
        Camera camera = scene.camera();
        if (camera.getProjectionMode() == se.llbit.chunky.renderer.projection.ProjectionMode.PINHOLE) {
            Vector3 pos = new Vector3(camera.getPosition());
            pos.sub(scene.getOrigin());
            float[] settings = new float[15];
            System.arraycopy(Util.vector3ToFloat(pos), 0, settings, 0, 3);
            System.arraycopy(Util.matrix3ToFloat(Reflection.getFieldValue(camera, "transform", Matrix3.class)), 0, settings, 3, 9);
            settings[12] = camera.infiniteDoF() ? 0 : (float) (camera.getSubjectDistance() / camera.getDof());
            settings[13] = (float) camera.getSubjectDistance();
            settings[14] = (float) Camera.clampedFovTan(camera.getFov());
            ctx.cameraSet(0, settings);
            return false;
        }
        int w = scene.width, h = scene.height;
        float[] rays = new float[w * h * 6];
        double halfWidth = w / (2.0 * h), invHeight = 1.0 / h;
        IntStream.range(0, w).parallel().forEach(i -> {
            Ray ray = new Ray();
            for (int j = 0; j < h; j++) {
                float ox = jitter ? ThreadLocalRandom.current().nextFloat() : 0.5f;
                float oy = jitter ? ThreadLocalRandom.current().nextFloat() : 0.5f;
                camera.calcViewRay(ray, -halfWidth + (i + ox) * invHeight, -0.5 + (j + oy) * invHeight);
                ray.o.sub(scene.getOrigin());
                int at = (j * w + i) * 6;
                System.arraycopy(Util.vector3ToFloat(ray.o), 0, rays, at, 3);
                System.arraycopy(Util.vector3ToFloat(ray.d), 0, rays, at + 3, 3);
            }
        });
        // The reference takes renderLock around the upload because its command queue is shared with the pass loop
        // (ClCamera.java:99-104).  ccu_camera_set uploads into the ray buffer no launch is reading, on the library's copy stream,
        // so the upload overlaps the passes in flight and the next ccu_render_passes picks the new rays up: no lock needed.
        ctx.cameraSet(-1, rays);
        return true;
    }
}
