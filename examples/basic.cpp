// The reference's examples/basic.rs (lines 22-33) through the C++ facade: prints the tiles and
// spans of one quadratic+cubic path.  Build: see INTEGRATION.md.  Needs a CUDA device.
#include <cstdio>

#include "../include/ochre.hpp"

using namespace ochre;

struct Printer : TileBuilder {
    void tile(int16_t x, int16_t y, const std::array<uint8_t, 64>& data) override {
        std::printf("tile at (%d, %d):\n", x, y);
        for (size_t row = 0; row < TILE_SIZE; ++row) {
            std::printf("  ");
            for (size_t col = 0; col < TILE_SIZE; ++col) std::printf("%3d ", data[row * TILE_SIZE + col]);
            std::printf("\n");
        }
    }
    void span(int16_t x, int16_t y, uint16_t width) override { std::printf("span at (%d, %d), width %d\n", x, y, width); }
};

int main() {
    Context ctx(0);
    Printer builder;
    Rasterizer rasterizer(ctx);
    rasterizer.fill({PathCmd::Move(Vec2(400, 300)), PathCmd::Quadratic(Vec2(500, 200), Vec2(400, 100)),
                     PathCmd::Cubic(Vec2(350, 150), Vec2(100, 250), Vec2(400, 300)), PathCmd::Close()},
                    Transform::id());
    rasterizer.finish(builder);

    // The same outline as a document of two paints -- a fill and a 3 px stroke -- in one submission: the stroke is
    // flattened and offset on the device (Rasterizer::stroke, reference src/rasterizer.rs:169-171).
    struct Counter : TileBuilder {
        size_t tiles = 0, spans = 0;
        void tile(int16_t, int16_t, const std::array<uint8_t, 64>&) override { ++tiles; }
        void span(int16_t, int16_t, uint16_t) override { ++spans; }
    } fill_count, stroke_count;
    const std::vector<PathCmd> outline = {PathCmd::Move(Vec2(400, 300)), PathCmd::Quadratic(Vec2(500, 200), Vec2(400, 100)),
                                          PathCmd::Cubic(Vec2(350, 150), Vec2(100, 250), Vec2(400, 300)), PathCmd::Close()};
    std::vector<Paint> paints = {Paint{outline, Transform::id(), 0.0f}, Paint{outline, Transform::id(), 3.0f}};
    std::vector<TileBuilder*> builders = {&fill_count, &stroke_count};
    finish_paints(ctx, paints, builders);
    std::printf("fill: %zu tiles, %zu spans; 3 px stroke: %zu tiles, %zu spans\n", fill_count.tiles, fill_count.spans, stroke_count.tiles,
                stroke_count.spans);
    return 0;
}
