package dev.thatredox.chunkynative.cuda;

import dev.thatredox.chunkynative.common.export.texture.AbstractTextureLoader;
import dev.thatredox.chunkynative.common.export.texture.TextureRecord;
import it.unimi.dsi.fastutil.objects.Object2ObjectMap;
import se.llbit.chunky.resources.Texture;

import java.util.ArrayList;
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
        // ClTextureLoader.java:32-44: stable sort by packed size, largest first; first fit scanning x outer / y inner over the layers
        // in order; a new layer only when no existing one takes the texture
        List<Placed> all = new ArrayList<>();
        textures.forEach((t, r) -> all.add(new Placed(t, r)));
        all.sort((a, b) -> b.size - a.size);

        List<boolean[][]> layers = new ArrayList<>();
        layers.add(new boolean[TILES][TILES]);
        for (Placed p : all) {
            if (!insert(layers, p)) {
                layers.add(new boolean[TILES][TILES]);
                insert(layers, p);
            }
        }

        ctx.atlasCreate(TILES * TILE, TILES * TILE, layers.size());
        for (Placed p : all) {
            ctx.atlasWrite(p.x * TILE, p.y * TILE, p.layer, p.width(), p.height(), rgba8(p.texture));
            p.record.set(((long) p.size << 32) | (((long) p.x << 22) | ((long) p.y << 13) | p.layer) & 0xFFFFFFFFL);   // :123-132
        }
    }

    /** ClTextureLoader.java:72-87: tile counts are width / 16 and height / 16, rounded DOWN, exactly as the reference reserves them. */
    private static boolean insert(List<boolean[][]> layers, Placed p) {
        int l = 0;
        for (boolean[][] layer : layers) {
            for (int x = 0; x < TILES; x++)
                for (int y = 0; y < TILES; y++)
                    if (insertAt(x, y, p.width() / TILE, p.height() / TILE, layer)) {
                        p.x = x; p.y = y; p.layer = l;
                        return true;
                    }
            l++;
        }
        return false;
    }

    /** ClTextureLoader.java:89-113 */
    private static boolean insertAt(int x, int y, int width, int height, boolean[][] layer) {
        if (y + height > layer.length || x + width > layer[0].length) return false;
        for (int line = y; line < y + height; line++)
            for (int pixel = x; pixel < x + width; pixel++)
                if (layer[line][pixel]) return false;
        for (int line = y; line < y + height; line++)
            for (int pixel = x; pixel < x + width; pixel++)
                layer[line][pixel] = true;
        return true;
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
