package dev.thatredox.chunkynative.cuda;

import dev.thatredox.chunkynative.common.export.texture.AbstractTextureLoader;
import dev.thatredox.chunkynative.common.export.texture.TextureRecord;
import it.unimi.dsi.fastutil.objects.Object2ObjectMap;
import se.llbit.chunky.resources.Texture;

import java.util.ArrayList;
import java.util.Comparator;
import java.util.List;

/**
 * Replacement for opencl.renderer.export.ClTextureLoader.  Placement is the reference's: textures sorted by packed size
 * (largest first), first fit on a 256 x 256 grid of 16-pixel tiles per 8192 x 8192 layer, record = size << 32 | x << 22 |
 * y << 13 | layer (ClTextureLoader.java:32-70,123-152) - the device decodes exactly those fields (textureAtlas.h:10-16).  The
 * sink changes: one ccu_scene_atlas_write per texture instead of clEnqueueWriteImage; the library stores the atlas
 * tile-linear.  NOT COMPILED in the build image.
 */
public class CudaTextureLoader extends AbstractTextureLoader {
    private static final int TILES = 256, TILE = 16;
    private final ChunkyCu.Context ctx;

    public CudaTextureLoader(ChunkyCu.Context ctx) { this.ctx = ctx; }

    private static final class Placed {
        final Texture texture; final TextureRecord record; final int size;
        int x, y, layer;
        Placed(Texture t, TextureRecord r) { texture = t; record = r; size = (t.getWidth() << 16) | t.getHeight(); }
        int width() { return size >>> 16; }
        int height() { return size & 0xFFFF; }
    }

    @Override
    protected void buildTextures(Object2ObjectMap<Texture, TextureRecord> textures) {
        List<Placed> all = new ArrayList<>();
        textures.forEach((t, r) -> all.add(new Placed(t, r)));
        all.sort(Comparator.comparingInt((Placed p) -> p.size).reversed());

        List<boolean[][]> used = new ArrayList<>();
        for (Placed p : all) {
            int w = (p.width() + TILE - 1) / TILE, h = (p.height() + TILE - 1) / TILE;
            boolean done = false;
            for (int layer = 0; !done; layer++) {
                if (layer == used.size()) used.add(new boolean[TILES][TILES]);
                boolean[][] grid = used.get(layer);
                for (int x = 0; x + w <= TILES && !done; x++)
                    for (int y = 0; y + h <= TILES && !done; y++)
                        if (free(grid, x, y, w, h)) {
                            take(grid, x, y, w, h);
                            p.x = x; p.y = y; p.layer = layer;
                            done = true;
                        }
            }
        }

        ctx.atlasCreate(TILES * TILE, TILES * TILE, Math.max(1, used.size()));
        for (Placed p : all) {
            ctx.atlasWrite(p.x * TILE, p.y * TILE, p.layer, p.width(), p.height(), rgba8(p.texture));
            p.record.set(((long) p.size << 32) | (((long) p.x << 22) | ((long) p.y << 13) | p.layer) & 0xFFFFFFFFL);
        }
    }

    private static boolean free(boolean[][] grid, int x, int y, int w, int h) {
        for (int j = y; j < y + h; j++) for (int i = x; i < x + w; i++) if (grid[j][i]) return false;
        return true;
    }

    private static void take(boolean[][] grid, int x, int y, int w, int h) {
        for (int j = y; j < y + h; j++) for (int i = x; i < x + w; i++) grid[j][i] = true;
    }

    /** Linear float colour -> byte, truncating, as ClTextureLoader.java:154-168. */
    private static byte[] rgba8(Texture t) {
        byte[] out = new byte[t.getWidth() * t.getHeight() * 4];
        int k = 0;
        for (int y = 0; y < t.getHeight(); y++)
            for (int x = 0; x < t.getWidth(); x++) {
                float[] c = t.getColor(x, y);
                for (int ch = 0; ch < 4; ch++) out[k++] = (byte) (c[ch] * 255.0);
            }
        return out;
    }
}
