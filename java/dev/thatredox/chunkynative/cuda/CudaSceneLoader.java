package dev.thatredox.chunkynative.cuda;

import dev.thatredox.chunkynative.common.export.AbstractSceneLoader;
import dev.thatredox.chunkynative.common.export.ResourcePalette;
import dev.thatredox.chunkynative.common.export.primitives.PackedBlock;
import se.llbit.chunky.renderer.scene.Scene;

import java.util.Arrays;

/**
 * Replacement for opencl.renderer.ClSceneLoader: the device-agnostic packers in common/export are reused unchanged;
 * only the sinks change from ClIntBuffer / cl_mem to ccu_scene_set_* calls.  Sketch - the palette factories
 * (createBlockPalette() etc.) return plain IntArrayList-backed palettes whose build() hands the int[] to the context.
 * NOT COMPILED in the build image.
 */
public abstract class CudaSceneLoader extends AbstractSceneLoader {
    protected final ChunkyCu.Context ctx;

    protected CudaSceneLoader(int deviceIndex) { this.ctx = new ChunkyCu.Context(deviceIndex); }

    public ChunkyCu.Context context() { return ctx; }

    @Override
    protected boolean loadOctree(int[] octree, int depth, int[] blockMapping, ResourcePalette<PackedBlock> blockPalette) {
        // same leaf remap as ClSceneLoader.java:56-58
        int[] mapped = Arrays.stream(octree).map(i -> i > 0 || -i >= blockMapping.length ? i : -blockMapping[-i]).toArray();
        ctx.setOctree(mapped, depth);
        ctx.sceneCommit();
        return true;
    }

    /** ClCamera.java:33-105: 15 floats for the pinhole projector, or 6 floats per pixel of pre-generated rays. */
    public abstract void uploadCamera(Scene scene);
}
