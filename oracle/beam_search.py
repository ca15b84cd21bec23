"""TEST INFRASTRUCTURE (oracle): T5X `decoding.beam_search` restated in numpy.

The reference decodes with `decode_fn=decoding.beam_search` (mt3/models.py:127) at T5X's default
num_decodes=1 (notebook: `predict_batch_with_aux(..., decoder_params={'decode_rng': None})`).
T5X is a third-party dependency that is not vendored, pinned or installable here (setup.py:39-56),
so this file restates the published algorithm (t5x/decoding.py `beam_search`, itself the Flax WMT
example's) from its documentation and the call site: **parity unpinned** -- there is no reference-held
vector for it.  What it pins instead is the DIFFERENCE between that algorithm at beam size 1 and the
greedy loop, on crafted logits, so that the divergence is a tested fact and not prose.

Algorithm (per batch element, K = num_decodes live beams, alpha = 0.6, NEG_INF = -1e7):
  every step the K live prefixes are extended by every vocabulary item, the 2K best extensions by
  cumulative log-probability are kept; of those, the ones that do NOT end in EOS compete for the K live
  slots, the ones that DO end in EOS compete (by log-prob / brevity_penalty(alpha, length)) with the
  finished hypotheses found so far for the K finished slots; brevity_penalty(alpha, n) = ((5 + n) / 6)^alpha.
  The search stops at max_decode_len or once, for every batch element, the worst kept finished score
  beats the best score any live prefix could still reach (live log-prob / brevity_penalty(alpha,
  max_decode_len)).  The result is the best finished hypothesis, or the best live prefix if none finished.

At K = 1 this is NOT greedy-until-EOS: when EOS ranks first, the runner-up token keeps the search
alive, and an EOS that only ranks SECOND also produces a finished hypothesis; the winner is the
hypothesis with the best length-normalised score, which can end earlier or later than greedy's.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np

NEG_INF = -1.0e7
EOS_ID = 1


def brevity_penalty(alpha: float, length) -> float:
    return np.power((5.0 + length) / 6.0, alpha)


def log_softmax(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def beam_search(logits_fn: Callable[[np.ndarray, int], np.ndarray], batch: int, max_decode_len: int, eos_id: int = EOS_ID,
                num_decodes: int = 1, alpha: float = 0.6) -> Tuple[np.ndarray, np.ndarray]:
    """logits_fn(prefixes int [batch, K, step], step) -> float [batch, K, V]: next-token logits of every live prefix.
    Returns (tokens int32 [batch, max_decode_len] -- the best hypothesis, 0-padded after its EOS -- and its score)."""
    K = num_decodes
    live = np.zeros((batch, K, max_decode_len), np.int64)
    live_lp = np.tile(np.array([0.0] + [NEG_INF] * (K - 1)), (batch, 1))
    fin = np.zeros((batch, K, max_decode_len), np.int64)
    fin_score = np.full((batch, K), NEG_INF)
    fin_flag = np.zeros((batch, K), bool)
    bp_max = brevity_penalty(alpha, max_decode_len)
    for i in range(max_decode_len):
        # loop condition of the reference: stop once no live prefix can still beat the worst kept finished hypothesis
        worst_fin = np.where(fin_flag, fin_score, NEG_INF).min(axis=1)
        if i > 0 and np.all(worst_fin > live_lp.max(axis=1) / bp_max):
            break
        logp = log_softmax(np.asarray(logits_fn(live[:, :, :i], i), np.float64)) + live_lp[:, :, None]
        V = logp.shape[-1]
        flat = logp.reshape(batch, K * V)
        top = np.argsort(-flat, axis=1, kind="stable")[:, :2 * K]              # 2K best extensions
        top_lp = np.take_along_axis(flat, top, axis=1)
        beam, tok = top // V, top % V
        seqs = np.take_along_axis(live, beam[:, :, None], axis=1)
        seqs[:, :, i] = tok
        newly = tok == eos_id
        # live: the K best extensions that did not just end
        alive_lp = np.where(newly, NEG_INF, top_lp)
        keep = np.argsort(-alive_lp, axis=1, kind="stable")[:, :K]
        live = np.take_along_axis(seqs, keep[:, :, None], axis=1)
        live_lp = np.take_along_axis(alive_lp, keep, axis=1)
        # finished: extensions that just ended compete with the kept finished ones on the length-normalised score
        cand = np.where(newly, top_lp / brevity_penalty(alpha, i + 1), NEG_INF)
        all_seq = np.concatenate([fin, seqs], axis=1)
        all_score = np.concatenate([fin_score, cand], axis=1)
        all_flag = np.concatenate([fin_flag, newly], axis=1)
        keep = np.argsort(-all_score, axis=1, kind="stable")[:, :K]
        fin = np.take_along_axis(all_seq, keep[:, :, None], axis=1)
        fin_score = np.take_along_axis(all_score, keep, axis=1)
        fin_flag = np.take_along_axis(all_flag, keep, axis=1)
    any_fin = fin_flag.any(axis=1)
    best_seq = np.where(any_fin[:, None], fin[:, 0], live[:, 0])
    best_score = np.where(any_fin, fin_score[:, 0], live_lp[:, 0])
    out = np.zeros((batch, max_decode_len), np.int32)
    for b in range(batch):
        row = best_seq[b]
        if any_fin[b]:
            end = int(np.argmax(row == eos_id))
            out[b, :end + 1] = row[:end + 1]
        else:
            out[b] = row
    return out, best_score


def greedy(logits_fn: Callable[[np.ndarray, int], np.ndarray], batch: int, max_decode_len: int, eos_id: int = EOS_ID) -> np.ndarray:
    """The greedy loop the CUDA path runs by default (mt3_generate): argmax each step, stop at the first EOS."""
    out = np.zeros((batch, max_decode_len), np.int32)
    done = np.zeros((batch,), bool)
    for i in range(max_decode_len):
        lg = np.asarray(logits_fn(out[:, None, :i].astype(np.int64), i))[:, 0]
        nxt = np.where(done, 0, lg.argmax(-1))
        out[:, i] = nxt
        done |= nxt == eos_id
        if done.all():
            break
    return out
