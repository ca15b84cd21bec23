"""Feature converter and model wrapper for an encoder-decoder with continuous inputs (the reference's models.py).

`ContinuousInputsEncDecFeatureConverter` (models.py:24-118) is the step between a task example and the network:
`inputs` float32 [frames, depth] is trimmed / zero-padded to the task's input length (0.0 rows -- a value in log-mel space,
not log(eps)), `targets` int32 to the target length, the decoder input is the target shifted right
behind a 0 and the loss weights mark the non-padding targets.  seqio's packing is a training feature and is refused.

`ContinuousInputsEncoderDecoderModel` (models.py:121-152) pairs a `network.Transformer` with the converter and a decode
function: `predict_batch` is T5X's predict_batch_with_aux at num_decodes = 1 (BOS 0, EOS 1, at most max_decode_length
steps) -- `decode_fn='beam_search'` is the reference's default (models.py:127) and runs the library's beam-size-1
bookkeeping, `decode_fn='greedy'` the plain argmax loop.  Training (loss, optimizer, label smoothing) is out of scope.
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Sequence

import numpy as np


class ContinuousInputsEncDecFeatureConverter:
    """Feature converter for an encoder-decoder with continuous inputs."""

    TASK_FEATURES = {'inputs': {'dtype': np.float32, 'rank': 2}, 'targets': {'dtype': np.int32, 'rank': 1}}
    MODEL_FEATURES = {'encoder_input_tokens': {'dtype': np.float32, 'rank': 2}, 'decoder_target_tokens': {'dtype': np.int32, 'rank': 1},
                      'decoder_input_tokens': {'dtype': np.int32, 'rank': 1}, 'decoder_loss_weights': {'dtype': np.int32, 'rank': 1}}

    def __init__(self, pack: bool = False):
        if pack:
            raise ValueError('packing is a training feature; the inference path converts one example per row')
        self.pack = False

    @staticmethod
    def _trim_or_pad(x: np.ndarray, length: int) -> np.ndarray:
        x = x[:length]
        if x.shape[0] == length:
            return x
        return np.concatenate([x, np.zeros((length - x.shape[0],) + x.shape[1:], x.dtype)], axis=0)

    def convert_example(self, features: Mapping[str, Any], task_feature_lengths: Mapping[str, int]) -> Dict[str, np.ndarray]:
        """One task example -> the model features of fixed length (models.py:77-98 after seqio's trim / pad)."""
        missing = [k for k in task_feature_lengths if k not in features]
        if missing:
            raise ValueError('task_feature_lengths names features the example does not have: %s' % missing)
        inputs = np.asarray(features['inputs'], np.float32)
        if inputs.ndim != 2:
            raise ValueError('inputs must have rank 2 (frames, depth); got shape %s' % (inputs.shape,))
        targets = np.asarray(features.get('targets', np.zeros((0,), np.int32)), np.int32)
        if targets.ndim != 1:
            raise ValueError('targets must have rank 1; got shape %s' % (targets.shape,))
        inputs = self._trim_or_pad(inputs, int(task_feature_lengths['inputs']))
        targets = self._trim_or_pad(targets, int(task_feature_lengths['targets']))
        shifted = np.concatenate([np.zeros((1,), np.int32), targets[:-1]]) if targets.size else targets
        return {'encoder_input_tokens': inputs, 'decoder_target_tokens': targets, 'decoder_input_tokens': shifted,
                'decoder_loss_weights': (targets != 0).astype(np.int32)}

    def __call__(self, ds: Sequence[Mapping[str, Any]], task_feature_lengths: Mapping[str, int]) -> List[Dict[str, np.ndarray]]:
        return [self.convert_example(ex, task_feature_lengths) for ex in ds]

    def get_model_feature_lengths(self, task_feature_lengths: Mapping[str, int]) -> Dict[str, int]:
        """Length relationship between task and model features (models.py:100-118)."""
        enc, dec = task_feature_lengths['inputs'], task_feature_lengths['targets']
        return {'encoder_input_tokens': enc, 'decoder_target_tokens': dec, 'decoder_input_tokens': dec, 'decoder_loss_weights': dec}


class ContinuousInputsEncoderDecoderModel:
    """Encoder-decoder model with continuous inputs: a `network.Transformer` (or anything with its `generate`) behind
    T5X's predict_batch surface."""

    FEATURE_CONVERTER_CLS = ContinuousInputsEncDecFeatureConverter
    DECODE_FNS = {'beam_search': 'beam1', 'greedy': 'greedy'}

    def __init__(self, module, input_vocabulary=None, output_vocabulary=None, optimizer_def=None, input_depth: int = 512,
                 decode_fn: str = 'beam_search', label_smoothing: float = 0.0, z_loss: float = 0.0, loss_normalizing_factor=None):
        if decode_fn not in self.DECODE_FNS:
            raise ValueError("decode_fn must be 'beam_search' (T5X decoding.beam_search at num_decodes=1) or 'greedy'")
        if optimizer_def is not None or label_smoothing or z_loss or loss_normalizing_factor is not None:
            raise ValueError('inference only: optimizer / loss arguments are not supported')
        self.module = module
        self.input_vocabulary, self.output_vocabulary = input_vocabulary, output_vocabulary
        self._input_depth = int(input_depth)
        self._decode = self.DECODE_FNS[decode_fn]

    def get_initial_variables(self, rng=None, input_shapes=None, input_types=None):
        """Shape check of the reference's override (models.py:140-152): a rank-2 encoder shape gets the input depth appended,
        a rank-3 one must end in it.  Returns the completed shapes (there are no variables to initialise here)."""
        del rng, input_types
        enc = tuple(input_shapes['encoder_input_tokens'])
        if len(enc) == 2:
            enc = (*enc, self._input_depth)
        else:
            assert enc[-1] == self._input_depth
        return {**dict(input_shapes), 'encoder_input_tokens': enc}

    def predict_batch_with_aux(self, params, batch: Mapping[str, Any], rng=None, decoder_params=None, num_decodes: int = 1,
                               return_all_decodes: bool = False):
        """tokens int32 [B, max_decode_length] (0 after EOS) and an empty aux dict.  `params` is accepted for signature
        fidelity: the weights live in the module."""
        del params, rng, return_all_decodes
        if num_decodes != 1:
            raise ValueError('only num_decodes=1 is built (the notebook and infer.gin use the default)')
        if decoder_params and decoder_params.get('decode_rng') is not None:
            raise ValueError('decoding is deterministic: decode_rng must be None')
        x = batch['encoder_input_tokens']
        if x.shape[-1] != self._input_depth:
            raise ValueError('encoder_input_tokens depth %d != input_depth %d' % (x.shape[-1], self._input_depth))
        return self.module.generate(x, stop_at_eos=True, decode=self._decode), {}

    def predict_batch(self, params, batch: Mapping[str, Any], rng=None, decoder_params=None):
        return self.predict_batch_with_aux(params, batch, rng=rng, decoder_params=decoder_params)[0]
