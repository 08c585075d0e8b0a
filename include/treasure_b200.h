/*
 * treasure_b200 -- C ABI of the B200-native batched Treasure Game simulator.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference
 * (sd-james/gym-treasure-game) is pure Python and has no FFI of its own; the
 * interface user code programs against is the Gym protocol of
 *   gym_treasure_game/envs/treasure_game.py:54-114  (TreasureGame.reset/step/
 *   available_mask/render) and :38-51 (ObservationWrapper),
 * backed by _treasure_game_impl/_treasure_game_impl.py:19-481 (_TreasureGameImpl)
 * and the option layer (_option.py:20-36, _move_options.py:16-460).  Each entry
 * point below names the reference function(s) it replaces; INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no CUDA or torch types (streams are
 *     passed as `void*` holding a cudaStream_t / CUstream; NULL = legacy stream).
 *   - Pointers marked DEV are device pointers on the env's device, owned by the
 *     caller; HOST pointers are host memory (pinned memory makes the *_host
 *     calls asynchronous up to their final synchronisation).
 *   - Every call returns 0 on success or a negative tg_status; the message is
 *     available from tg_last_error() (thread-local).  No exceptions, no aborts.
 *   - A tg_env is bound to one device and is not thread-safe; distinct tg_env
 *     objects are independent.  All device work is enqueued on the given stream
 *     and is asynchronous unless stated otherwise.
 *   - There is NO CPU fallback: every compute entry point fails with
 *     TG_ERR_CUDA when no CUDA device is usable.
 */
#ifndef TREASURE_B200_H
#define TREASURE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_ABI_VERSION 1

/* compile-time limits of one level (the shipped level uses 3/2/1/2 objects on a 14x13 grid) */
#define TG_MAX_DOORS   6
#define TG_MAX_HANDLES 4
#define TG_MAX_BOLTS   3
#define TG_MAX_ITEMS   4      /* keys + gold coins */
#define TG_MAX_OBJECTS 16
#define TG_MAX_TRIGGERS 32
#define TG_MAX_GRID    26     /* cells per side */
#define TG_MAX_LEVELS  8      /* distinct layouts in one batch */
#define TG_NUM_OPTIONS 9      /* _treasure_game_impl.py:495 */
#define TG_TICK_CAP    4096   /* primitive ticks per option before the env is flagged (reference would hang) */

#define TG_FRAME_CHANNELS 3
#define TG_CELL_PX 48         /* _scale.py:8-9 */

typedef enum {
    TG_OK = 0,
    TG_ERR_ARG = -1,          /* bad argument / level outside the limits above */
    TG_ERR_CUDA = -2,         /* CUDA runtime error or no device */
    TG_ERR_STATE = -3,        /* call not valid in this state (e.g. render without sprites) */
    TG_ERR_NOMEM = -4
} tg_status;

/* object kinds, in the spelling of domain-objects.txt (_treasure_game_impl.py:127-163) */
enum { TG_DOOR = 0, TG_HANDLE = 1, TG_KEY = 2, TG_BOLT = 3, TG_GOLD = 4 };

/* option ids == gym action ids (_treasure_game_impl.py:495) */
enum { TG_GO_LEFT = 0, TG_GO_RIGHT, TG_UP_LADDER, TG_DOWN_LADDER, TG_INTERACT,
       TG_DOWN_LEFT, TG_DOWN_RIGHT, TG_JUMP_LEFT, TG_JUMP_RIGHT };

/* done byte written by tg_step */
#define TG_DONE_TERMINATED 1  /* gold in bag and player in row 0 (treasure_game.py:95) */
#define TG_DONE_TRUNCATED  2  /* max_episode_steps reached (no reference counterpart) */

typedef struct { int32_t kind, cx, cy, flag; } tg_object;          /* flag: closed / up / locked */
typedef struct { int32_t src_kind, src_index, src_value,
                         dst_kind, dst_index, dst_value; } tg_trigger; /* one line of domain-interactions.txt */

typedef struct tg_level tg_level;
typedef struct tg_env tg_env;

typedef struct {
    int32_t cw, ch;                     /* grid size in cells */
    int32_t n_doors, n_handles, n_bolts, n_items, n_objects, n_triggers;
    int32_t obs_dim;                    /* 2 + handles + bolts + 2*items (impl:368-378) */
    int32_t start_cx, start_cy;         /* first non-wall cell, row-major (impl:173-176) */
    int32_t frame_w, frame_h;           /* 48*cw, 48*ch */
    int32_t has_sprites;
} tg_level_info;

/* Flat, unpacked view of the per-env state; every pointer is DEV and may be NULL.
 * Strides are the TG_MAX_* constants regardless of the level. */
typedef struct {
    int32_t *pos;       /* [N][2]  playerx, playery                                  */
    int32_t *misc;      /* [N][4]  facing_right, jump_ticker, total_actions, draws   */
    uint8_t *doors;     /* [N][TG_MAX_DOORS]    closed                               */
    uint8_t *handles;   /* [N][TG_MAX_HANDLES]  up                                   */
    uint8_t *bolts;     /* [N][TG_MAX_BOLTS]    locked                               */
    double  *angles;    /* [N][TG_MAX_HANDLES]                                       */
    int32_t *items;     /* [N][TG_MAX_ITEMS][2] x, y in pixels                       */
    int32_t *bag;       /* [N][TG_MAX_ITEMS]    item index per bag slot, -1 = empty  */
    int64_t *acct;      /* [N][3]  episode return, episode steps, error flag | sticky handle flags << 1 */
} tg_state_view;

const char *tg_last_error(void);
int tg_abi_version(void);
/* number of usable CUDA devices (0 when there is none); never fails */
int tg_device_count(void);

/* ---- level: replaces the parsers' output (impl:180-202 tiles, impl:119-166 objects,
 *      impl:75-117 triggers).  HOST data, copied.  tiles: ch rows of cw chars, ' ' '/' 'L'. */
int tg_level_create(const uint8_t *tiles, int32_t cw, int32_t ch,
                    const tg_object *objs, int32_t n_objs,
                    const tg_trigger *trigs, int32_t n_trigs, tg_level **out);
int tg_level_get_info(const tg_level *lv, tg_level_info *out);
/* Render assets (replaces _treasure_game_drawer.py:59-134 sprite loading and the
 * static tile layer of draw_domain :140-152).  HOST data, copied.
 *   background_rgb : frame_h*frame_w*3 bytes, the precomposed tile layer
 *   sprites_rgba   : TG_NUM_SPRITES sprites of 48*48*4 bytes (order: tg_sprite_id) */
enum { TG_SPR_DOOR_CLOSED = 0, TG_SPR_DOOR_OPEN, TG_SPR_KEY, TG_SPR_GOLD, TG_SPR_BOLT_OPEN,
       TG_SPR_BOLT_LOCKED, TG_SPR_HANDLE_BASE, TG_SPR_HERO_RIGHT, TG_SPR_HERO_LEFT, TG_NUM_SPRITES };
int tg_level_set_sprites(tg_level *lv, const uint8_t *sprites_rgba, const uint8_t *background_rgb);
void tg_level_destroy(tg_level *lv);

/* ---- batch of environments: replaces TreasureGame.__init__ (treasure_game.py:63-76).
 * level_ids: HOST, num_envs bytes selecting levels[level_ids[i]], or NULL (all level 0).
 * Env i uses Philox stream (seed, first_env_id + i) so results do not depend on how a
 * population is sharded over devices.  max_episode_steps 0 = no time limit.
 * The envs are created reset (constructor draws, impl:31-53). */
int tg_create(const tg_level *const *levels, int32_t n_levels, const uint8_t *level_ids,
              int64_t num_envs, int64_t first_env_id, int32_t device, uint64_t seed,
              int32_t max_episode_steps, int32_t auto_reset, tg_env **out);
void tg_destroy(tg_env *env);
int64_t tg_num_envs(const tg_env *env);
int32_t tg_obs_dim(const tg_env *env);      /* max over the batch's levels */

/* TreasureGame.reset (treasure_game.py:78-81 -> impl:55-73).  mask DEV [N] (non-zero = reset) or NULL = all.
 * obs DEV [N][obs_dim] float32 or NULL. */
int tg_reset(tg_env *env, const uint8_t *mask, float *obs, void *stream);

/* TreasureGame.step (treasure_game.py:91-96 -> _option.py:20-36 -> impl:290-359).
 *   actions DEV [N] int32 option ids; obs DEV [N][obs_dim] f32; reward DEV [N] f32 (0 where the
 *   option was not runnable -- the reference returns None there, see `ran`); done DEV [N] u8
 *   (TG_DONE_* bits); ran DEV [N] u8 or NULL; avail DEV [N] u16 or NULL: 9-bit available mask of
 *   the state *after* the step (bit k = option k runnable; treasure_game.py:83-89).
 * With auto_reset, an env whose episode ended is reset inside the call and obs holds the first
 * observation of the new episode.
 * Observation rows: an option that cannot run leaves its env untouched (_option.py:22-23), so its row does not
 * change.  When `obs` is the buffer registered with tg_bind_obs and it was the `obs` argument of the previous
 * state-changing call on this env (tg_step / tg_reset / tg_primitive_step), only what changed is written -- the two
 * player slots of an env whose option moved the player and nothing else, the whole row of an env that interacted,
 * picked something up or was reset -- in one kernel launch.  Any other `obs` has every row written (a second launch). */
int tg_step(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done,
            uint8_t *ran, uint16_t *avail, void *stream);

/* Declares `obs` (DEV [N][obs_dim] f32) a persistent buffer of the caller that only this library writes: tg_step
 * calls that pass it then update just the rows that changed (see tg_step).  NULL unbinds. */
int tg_bind_obs(tg_env *env, float *obs);

/* Gym-style flag views: three persistent DEV u8 [N] arrays (4-byte aligned) that every later tg_step also fills with
 * 0 / 1 bytes -- done != 0, terminated, truncated -- next to the `done` bit mask, so that a host framework can hand them
 * out as boolean tensors without elementwise kernels of its own.  All three NULL unbinds.  (tg_step_host* and
 * tg_primitive_step do not write them.) */
int tg_bind_flags(tg_env *env, uint8_t *done01, uint8_t *terminated01, uint8_t *truncated01);

/* Same, with HOST buffers (pinned recommended): copies actions in, runs the step, copies
 * obs/reward/done(/ran) out and synchronises the stream.  Any output may be NULL. */
int tg_step_host(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                 uint8_t *ran, void *stream);

/* Same results in the same HOST arrays, but only what changed crosses the bus: an option that cannot run leaves its
 * env untouched (_option.py:22-23), so under the previous call's arrays only the envs that ran or were reset need
 * new rows.  The kernel compacts those into records on the device; the call copies the records and patches the
 * arrays (host threads).  Contract: obs / reward / done / ran are the arrays the previous tg_step_host or
 * tg_step_host_sparse call of this env filled, unmodified in `obs`; otherwise (first call, other pointers, or any
 * other state-changing call in between) the call falls back to tg_step_host.  All four outputs are required. */
int tg_step_host_sparse(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                        uint8_t *ran, void *stream);

/* tg_step_host_sparse in two halves, for callers that keep two (or more) batches in flight -- the double-buffered
 * vector env of an actor loop: while the host patches batch A's arrays, batch B's kernels run and its records cross.
 * _begin enqueues the action copy, the step kernels and the record copies and returns without waiting (a first call /
 * other arrays run tg_step_host inside _begin instead); _end waits for the records chunk by chunk and patches the
 * arrays given to _begin.  Contract between the two calls: `actions` and the four output arrays stay untouched (the
 * outputs still hold the previous step's results), no other state-changing call on this env (they return
 * TG_ERR_STATE), and each env in flight has its own `stream`.  tg_step_host_sparse == _begin followed by _end. */
int tg_step_host_sparse_begin(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                              uint8_t *ran, void *stream);
int tg_step_host_sparse_end(tg_env *env);

/* Debug: seconds the sparse host step has spent enqueueing (_begin), waiting for records and patching (_end) since the
 * last call of this function (then reset).  out3 HOST double[3]. */
void tg_debug_host_times(tg_env *env, double *out3);

/* Debug / tests: one entry of a level's trigger closure table (the tabulated form of set_val + process_trigger,
 * objects.py:76-94, :145-149).  obj = object index in file order, value = the value it is set to, bits = door bits (6) |
 * handle bits << 6 (4) | bolt bits << 10 (3) before the call.  *entry = bits afterwards | events << 13 | (handle index |
 * value << 2) << (16 + 3 k) for the k-th handle whose angle is redrawn; events == 7: not tabulated.  Needs no device. */
int tg_debug_level_closure(const tg_level *level, int32_t obj, int32_t value, uint32_t bits, uint32_t *entry);

/* TreasureGame.available_mask (treasure_game.py:83-89).  mask DEV [N][9] u8. */
int tg_available_mask(tg_env *env, uint8_t *mask, void *stream);

/* TreasureGame.render('rgb_array') (treasure_game.py:98-104 -> drawer.draw_domain :136-163).
 * frames DEV [count][frame_h][frame_w][3] u8 for envs first .. first+count-1. */
int tg_render(tg_env *env, int64_t first, int64_t count, uint8_t *frames, void *stream);

/* The option layer with a drawer (_option.py:20-36 with :33-34: drawer.draw_domain() after every primitive tick;
 * create_options(md, drawer), _move_options.py).  The `count` envs env_ids[k] (DEV int64, distinct) take one gym step
 * with option actions[k] (DEV int32), run tick by tick; frames DEV [count][max_ticks][frame_h][frame_w][3] u8
 * (16-byte aligned) receives the frame after tick t at [k][t-1]; frames after the option's last tick repeat its final
 * frame (an option that cannot run yields max_ticks copies of the unchanged state; ticks beyond max_ticks are run but
 * not drawn).  n_ticks DEV [count] int32 = primitive ticks of the option (0 = not runnable, -1 = bad env id).  obs DEV
 * [count][obs_dim] / reward / done / ran DEV [count], any of them NULL: the step's results as in tg_step (after an
 * auto-reset obs is the first observation of the new episode while the frames show the old one).  The other envs of
 * the batch are not stepped. */
int tg_step_frames(tg_env *env, const int64_t *env_ids, int32_t count, const int32_t *actions, int32_t max_ticks,
                   uint8_t *frames, int32_t *n_ticks, float *obs, float *reward, uint8_t *done, uint8_t *ran, void *stream);

/* ---- analysis helpers of the drawer (off the gym path; _treasure_game_drawer.py:165-231) ----
 * _TreasureGameDrawer.blend (:207-231), batched: surface s (s < count / per_surface) receives the envs
 * first + s*per_surface .. first + (s+1)*per_surface - 1 one after the other, each as blend(surf, alpha_objs,
 * alpha_player) would: objects except the handle bases on a transparent overlay blended with opacity alpha_objs,
 * handle bases at full opacity, the hero with opacity alpha_player.  per_surface = 1 is one blend() per env;
 * per_surface = count accumulates a set of states on one surface (the use the reference's research code makes
 * of it).  surfaces DEV [count / per_surface][frame_h][frame_w][3] u8, in/out, 16-byte aligned.  Opacities are
 * the reference's int(255 * alpha), 0..255. */
int tg_blend(tg_env *env, int64_t first, int64_t count, int64_t per_surface, uint8_t *surfaces,
             int32_t alpha_objs, int32_t alpha_player, void *stream);
/* _TreasureGameDrawer.draw_background_to_surface (:165-182): the tile layer of a level of the batch.
 * frame DEV [frame_h][frame_w][3] u8.  (draw_to_surface :184-196 has the pixels of tg_render.) */
int tg_background(tg_env *env, int32_t level, uint8_t *frame, void *stream);
/* _TreasureGameDrawer.blit_alpha (:198-205): source (RGB or RGBA, channels = 3 | 4) is blended over the region of
 * target at (x, y) and the result is put back with per-surface opacity.  target DEV [th][tw][3] u8 in/out,
 * source DEV [sh][sw][channels] u8; clipped to the target.  No tg_env: runs on the current device. */
int tg_blit_alpha(uint8_t *target, int32_t tw, int32_t th, const uint8_t *source, int32_t sw, int32_t sh,
                  int32_t channels, int32_t x, int32_t y, int32_t opacity, void *stream);

/* State save / restore on the SoA (test hook; superset of impl:368-378 / :447-481). */
int tg_get_state(tg_env *env, const tg_state_view *out, void *stream);
int tg_set_state(tg_env *env, const tg_state_view *in, void *stream);

/* _TreasureGameImpl.step(action) (_treasure_game_impl.py:290-359): one primitive action per env, no option layer.
 * actions DEV [N] int32 in _actions.py:7-13 (0 NOP, 1 UP, 2 DOWN, 3 LEFT, 4 RIGHT, 5 JUMP, 6 INTERACT; other ids act
 * as NOP like the reference's if-chain); reward -1, JUMP -5; obs / done / accounting / auto-reset as in tg_step. */
enum { TG_ACT_NOP = 0, TG_ACT_UP, TG_ACT_DOWN, TG_ACT_LEFT, TG_ACT_RIGHT, TG_ACT_JUMP, TG_ACT_INTERACT };
int tg_primitive_step(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done, void *stream);

/* _TreasureGameImpl.init_with_state (_treasure_game_impl.py:447-481), quirks included: states DEV [N][obs_dim]
 * float64 normalised state vectors in get_state_descriptors order (:380-400), -99 = keep the current value;
 * every key / gold / bolt reads the first slot of its name; facing is forced right; the bag, the jump ticker and
 * the counters are untouched; handle angles propagate their triggers (targets redraw their angle from the env's
 * RNG stream) and the handles stay flagged previously_triggered.  mask DEV [N] or NULL = all. */
int tg_init_with_state(tg_env *env, const double *states, const uint8_t *mask, void *stream);

/* Parity mode: uniform draws come from tape[offsets[i] + draw_index_i] instead of Philox.
 * Both DEV, must stay valid until replaced; (NULL, NULL) returns to Philox.  Resets draw indices to 0. */
int tg_set_draw_tape(tg_env *env, const double *tape, const int64_t *offsets, void *stream);

/* Episode statistics accumulated on the device since creation (or the last tg_stats_clear):
 * out8 DEV int64[8] = episodes, successes, sum return, sum episode steps, primitive ticks,
 * runnable steps, gym steps, errors.  This is the vector the host all-reduces over ranks. */
int tg_stats(tg_env *env, int64_t *out8, void *stream);
int tg_stats_clear(tg_env *env, void *stream);

/* Debug instrumentation: DEV uint64[8192 * 8 + 8192 * 512] (or NULL to switch off).  When set, thread 0 of every
 * step-kernel CTA (up to 8192) writes %globaltimer (ns) at its phase boundaries into stamps[cta][0..7]: start, levels
 * staged, phase A done, sorted, its warp left phase B, phase B done, phase C done, statistics done; and lane 0 of the warp
 * that ran chunk q of the CTA writes, at stamps[8192 * 8 + (cta * 128 + q) * 4 ..]: start time, duration | class << 32 |
 * ticks << 40, and the durations of the chunk's four segments (load + set-up, option, state + plan, observation +
 * outputs; two 32-bit values per word).  Used by tools/bench_phases.py and tools/bench_chunks.py; not part of the
 * reference surface. */
int tg_debug_phase_buffer(tg_env *env, uint64_t *stamps);

/* Debug / tuning: environments per step-kernel CTA (32..4096, rounded up to a multiple of 4; 0 = automatic, the
 * default, which also honours the TG_STEP_TILE environment variable read by tg_create).  Results never depend on it;
 * the parity tests use it to run every tile-size-dependent code path against the oracle. */
int tg_debug_set_step_tile(tg_env *env, int32_t tile);

/* how many kernels this library has launched on behalf of `env` (bench bookkeeping) */
int64_t tg_launch_count(const tg_env *env);
/* bytes the tg_step_host* calls of `env` have copied host->device / device->host so far (bench bookkeeping) */
void tg_host_traffic(const tg_env *env, int64_t *h2d_bytes, int64_t *d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* TREASURE_B200_H */
