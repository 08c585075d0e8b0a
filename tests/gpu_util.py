"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

import c_oracle
import py_oracle as po
from gym_treasure_game_b200 import Level


def product_level(lvt: "po.LevelText") -> Level:
    """The product's Level built through the product's own text front end."""
    return Level.from_strings(*po.level_to_strings(lvt))


def assert_state_equal(env, cb, msg=""):
    """Full per-env state of the CUDA batch vs the C oracle batch (bit-exact)."""
    st = {k: v.cpu().numpy() for k, v in env.get_state().items()}
    cs = cb.state()
    lv = cb.level
    np.testing.assert_array_equal(st["pos"], cs["pos"], err_msg="pos " + msg)
    np.testing.assert_array_equal(st["misc"], cs["misc"], err_msg="facing/ticker/total_actions/draws " + msg)
    np.testing.assert_array_equal(st["doors"][:, : lv.nd], cs["doors"], err_msg="doors " + msg)
    np.testing.assert_array_equal(st["handles"][:, : lv.nh], cs["handles"], err_msg="handles " + msg)
    np.testing.assert_array_equal(st["bolts"][:, : lv.nb], cs["bolts"], err_msg="bolts " + msg)
    np.testing.assert_array_equal(st["angles"][:, : lv.nh], cs["angles"], err_msg="angles (f64 bit-exact) " + msg)
    np.testing.assert_array_equal(st["items"][:, : lv.ni], cs["items"][:, :, :2], err_msg="items " + msg)
    np.testing.assert_array_equal(st["bag"], cs["bag"], err_msg="bag (ordered, duplicates allowed) " + msg)
    acct = st["acct"].copy()
    sticky = acct[:, 2] >> 1
    acct[:, 2] &= 1
    np.testing.assert_array_equal(acct, cs["acct"], err_msg="episode accounting " + msg)
    pt = cb.handles_pt()
    np.testing.assert_array_equal(np.stack([(sticky >> h) & 1 for h in range(lv.nh)], axis=1) if lv.nh else pt, pt,
                                  err_msg="sticky previously_triggered flags " + msg)


def assert_step_equal(out, ref, msg=""):
    obs, rew, done_bits, ran = out
    o2, r2, d2, ran2, _ = ref
    np.testing.assert_array_equal(rew.cpu().numpy(), r2, err_msg="reward " + msg)
    np.testing.assert_array_equal(done_bits.cpu().numpy(), d2, err_msg="done " + msg)
    np.testing.assert_array_equal(ran.cpu().numpy(), ran2, err_msg="ran " + msg)
    np.testing.assert_array_equal(obs.cpu().numpy()[:, : o2.shape[1]], o2.astype(np.float32), err_msg="obs " + msg)
