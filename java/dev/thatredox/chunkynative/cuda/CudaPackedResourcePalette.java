package dev.thatredox.chunkynative.cuda;

import dev.thatredox.chunkynative.common.export.Packer;
import dev.thatredox.chunkynative.common.export.ResourcePalette;
import it.unimi.dsi.fastutil.ints.IntArrayList;

/**
 * Replacement for opencl.renderer.export.ClPackedResourcePalette: the packed words are collected on the host and handed to
 * the context by CudaSceneLoader once the scene is complete (ccu_scene_set_*), instead of becoming a cl_mem on first use.
 * put() returns the word offset of the resource, exactly as the reference does (ClPackedResourcePalette.java:14-19), so every
 * pointer the common/export packers embed stays valid.  NOT COMPILED in the build image.
 */
public class CudaPackedResourcePalette<T extends Packer> implements ResourcePalette<T> {
    private final IntArrayList words = new IntArrayList();
    private boolean locked = false;

    @Override
    public int put(T resource) {
        if (locked) throw new IllegalStateException("Attempted to modify a locked palette.");
        int pointer = words.size();
        words.addAll(resource.pack());
        return pointer;
    }

    /** Backing array and its fill level; the palette is read-only from here on. */
    public int[] elements() { locked = true; return words.elements(); }
    public int size() { return words.size(); }
}
