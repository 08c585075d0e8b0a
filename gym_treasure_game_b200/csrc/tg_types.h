// Internal POD shared by host (level compiler in tg_capi.cu) and device code.
#pragma once
#include <cstddef>
#include <stdint.h>
#include <vector_types.h>
#include "../../include/treasure_b200.h"

namespace tg {

constexpr int S = TG_CELL_PX;          // _scale.py:8-9
constexpr int PAD = 3;                 // border of WALL cells around the grid in the tile table
constexpr int TSTRIDE = 32;            // tile table is 32 x 32 bytes

// effective cell types (after the door override, objs.py:246-253)
enum : int { T_OPEN = 0, T_WALL = 1, T_LADDER = 2, T_DOOR = 3 };
// tile-table byte: bits 0-1 base type, bit 2 static object (handle/bolt) in the cell,
// bit 3 a door lives here, bits 4-6 its door index
constexpr int TC_STATIC_OBJ = 4, TC_HAS_DOOR = 8;

// primitive actions, _actions.py:7-13
enum : int { A_NOP = 0, A_UP, A_DOWN, A_LEFT, A_RIGHT, A_JUMP, A_INTERACT };

// flags word layout (per env)
constexpr int F_FACING = 0;                       // 1 bit
constexpr int F_TICKER = 1;                       // 5 bits
constexpr int F_ERROR = 6;                        // 1 bit
constexpr int F_DOORS = 7;                        // TG_MAX_DOORS bits, 1 = closed
constexpr int F_HANDLES = F_DOORS + TG_MAX_DOORS; // 13, TG_MAX_HANDLES bits, 1 = up
constexpr int F_BOLTS = F_HANDLES + TG_MAX_HANDLES; // 17, TG_MAX_BOLTS bits, 1 = locked
// the bag is an ordered list that may hold an item twice (the reference appends on every pickup, impl:350-354,
// and init_with_state can put a bagged item back on the map): length + TG_MAX_ITEMS slots of 2 bits
constexpr int F_BAGLEN = F_BOLTS + TG_MAX_BOLTS;  // 20, 3 bits: entries in the bag (0..TG_MAX_ITEMS)
constexpr int F_SPARE = F_BAGLEN + 3;             // 23, 1 bit unused
constexpr int F_BAGORD = F_SPARE + 1;             // 24, TG_MAX_ITEMS x 2 bits: item index at bag position j
static_assert(F_BAGORD + 2 * TG_MAX_ITEMS == 32, "flags word is exactly 32 bits");

// One level, as staged in shared memory.  sizeof is a multiple of 16 (cp.async.bulk).
struct alignas(16) LevelBlob {
    uint8_t tiles[TSTRIDE * TSTRIDE];   // [(cy+PAD)*32 + (cx+PAD)]
    int16_t cw, ch;
    int16_t start_px, start_py;          // centre-x / top-y of the start cell before noise (impl:176)
    uint8_t n_doors, n_handles, n_bolts, n_items, n_objs, n_trigs, obs_dim, key_mask; // key_mask: bit i = item i is a key
    uint32_t init_flags;                 // facing=1, ticker=0, object bits from the level file
    uint32_t gold_mask;                  // bit i = item i is gold
    uint8_t obj_kind[TG_MAX_OBJECTS];    // file order (impl:127-163)
    uint8_t obj_idx[TG_MAX_OBJECTS];     // index within its kind
    uint8_t obj_obs[TG_MAX_OBJECTS];     // first obs slot of the object (impl:368-378), 255 = none
    int8_t door_cx[8], door_cy[8];
    int8_t handle_cx[4], handle_cy[4];
    int8_t bolt_cx[4], bolt_cy[4];
    int8_t item_cx[4], item_cy[4];       // initial cells
    uint8_t handle_obj[4], bolt_obj[4], item_obj[4], door_obj[8];   // object (file-order) index of each
    uint8_t trig_src[TG_MAX_TRIGGERS];   // object index | value << 7, file order (objs.py:65-71)
    uint8_t trig_dst[TG_MAX_TRIGGERS];
    // the same table grouped by source: the triggers registered for (object o, value v) are trig_list[trig_begin[2o+v] ..
    // trig_begin[2o+v+1]) in file order, each entry the target as object index | value << 7 (process_trigger walks only these)
    uint8_t trig_begin[2 * TG_MAX_OBJECTS + 8];
    uint8_t trig_list[TG_MAX_TRIGGERS];
    float inv_w, inv_h;                  // unused by parity paths (obs divides in double)
    // closure of set_val + process_trigger (objs:76-94, :145-149) for an env without sticky handles, in GLOBAL memory (or
    // NULL): entry [(2 o + v) << 13 | door / handle / bolt bits] = new bits | events << 13 | (handle | value << 2) << (16 + 3 k)
    // for the k-th handle whose angle is redrawn, in the walk's order; events == 7: not tabulated, walk the graph
    const uint32_t *closure;
    // observation program (impl:368-378): obs slot k = OP_* << 4 | index (handle / bolt / item number)
    uint8_t obs_prog[32];
    // row bit masks over the padded columns (bit = column index + PAD), door cells cleared:
    uint32_t row_nonopen[TSTRIDE];       // WALL or LADDER (anything but OPEN)          -> can_fall, up_clear
    uint32_t row_solid[TSTRIDE];         // WALL                                          -> can_go_left/right
    uint32_t row_ladder[TSTRIDE];        // LADDER                                        -> can_go_up/down
    uint32_t row_static_obj[TSTRIDE];    // a handle or bolt lives in the cell (is_object_at, impl:402-409)
    uint32_t col_nonopen[TSTRIDE];       // column profiles of the same tables (bit = padded row): the ladder options never
    uint32_t col_ladder[TSTRIDE];        //   change playerx, so their probes look at fixed columns (door cells cleared)
    uint8_t row_lut[TSTRIDE];            // 0 = no door in this row, else 1 + index into door_lut
    uint32_t door_lut[TG_MAX_DOORS][64]; // [rows with doors][closed-door bits] -> column bits of the closed doors of the row
};
static_assert(sizeof(LevelBlob) % 16 == 0, "LevelBlob must be a multiple of 16 bytes");
static_assert(offsetof(LevelBlob, closure) % 8 == 0, "LevelBlob::closure must be 8-byte aligned");
constexpr int CLOSURE_BITS = TG_MAX_DOORS + TG_MAX_HANDLES + TG_MAX_BOLTS;   // 13: the flag bits from F_DOORS up

// observation program opcodes and the quotient tables behind OP_PX .. OP_IY: lut[0][v + S] = float(v / W) and
// lut[1][v + S] = float(v / H) for v in [-S, OBS_LUT_N - S), one pair of tables per level (BatchView::obs_lut)
enum : int { OP_PX = 0, OP_PY, OP_ANGLE, OP_BOLT, OP_IX, OP_IY, OP_ZERO };
constexpr int OBS_LUT_N = TG_MAX_GRID * TG_CELL_PX + TG_CELL_PX;

// Plan word per env (BatchView::plan): everything the step kernel needs to know about an env whose option does not
// run, derived from the env's state whenever that state changes (tg_device.cuh compute_plan), so that an idle env
// costs the step kernel one 8-byte load instead of its 16-byte state record and a 9-way can_run switch:
//   bits 0-8   can_run of option k (tg:83-89 available_mask)
//   bit  9     done: gold in the bag and the player in row 0 (tg:95)
//   bits 10-11 down_left / down_right would raise in the reference (target row None, opts:213-221)
//   bits 16+4k estimated length class of option k (0..11, estimate / 10 ticks): sort key only, any value is correct
constexpr int PL_TERM = 9, PL_ERR_DL = 10, PL_ERR_DR = 11, PL_LEN = 16;

// stats vector slots (tg_stats)
enum : int { ST_EPISODES = 0, ST_SUCCESS, ST_RETURN, ST_EPSTEPS, ST_TICKS, ST_RAN, ST_STEPS, ST_ERRORS };

// device-side description of one batch (passed by value to kernels)
struct BatchView {
    int64_t n;
    int64_t first_env_id;
    int64_t r_begin, r_count;   // env range one step launch covers ([0, n) unless tg_step_host pipelines chunks)
    uint4 *core;          // [N]  x=pos(px | py<<16)  y=flags  z=item0 (x | y<<16)  w=item1
    uint4 *acct;          // [N]  x=draws  y=ep_return  z=spare  w=total_actions
    uint64_t *plan;       // [N]  see PL_* above
    uint32_t *ep_start;   // [N]  index (value of *step_counter) of the first gym step of the env's current episode:
                          //      episode length so far = *step_counter - ep_start, no per-step write for idle envs
    uint32_t *step_counter;   // [2] device words: [0] gym steps every env of the batch has taken (one per tg_step /
                              //     tg_primitive_step call), [1] CTA ticket of the running launch
    int32_t advance;      // this launch is the last one of its API call: its last CTA increments step_counter[0]
    uint2 *items23;       // [N]  items 2,3 (NULL when every level has <= 2 items)
    double *angles;       // [TG_MAX_HANDLES][N]
    const uint8_t *level_id;   // [N] or NULL
    const LevelBlob *levels;   // [n_levels] in global memory
    int32_t n_levels;
    int32_t obs_dim;      // row stride of obs
    int32_t max_steps;
    int32_t auto_reset;
    uint32_t seed_lo, seed_hi;
    const double *tape;   // parity mode, or NULL
    const int64_t *tape_off;
    unsigned long long *stats;   // [8]
    const float *obs_lut;        // [n_levels][2][OBS_LUT_N]
    // sparse outputs (tg_step_host_sparse), or sp_count == NULL: one record of sp_words 32-bit words per env whose
    // outputs are not (obs unchanged, reward 0, done 0, ran 0): [0] env index, [1] reward (float bits),
    // [2] done | ran << 8, [3 ..] observation.  sp_count[0] = records of the launch (atomic); sp_count[4 + 2 b], [5 + 2 b] =
    // first record and number of records of CTA b's tile (its records are contiguous)
    uint32_t *sp_count;
    uint32_t *sp_recs;
    int32_t sp_words;
    // optional 0 / 1 byte views of the step's done bits (tg_bind_flags): done != 0, terminated, truncated -- what a Gym-style
    // caller wants as bool tensors, written by the step kernel itself instead of three elementwise launches afterwards
    uint8_t *flag_done, *flag_term, *flag_trunc;
    unsigned long long *phase_ts; // debug: [grid][8] globaltimer stamps of the step kernel's phases, or NULL
};

}  // namespace tg
