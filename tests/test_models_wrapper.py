"""mt3_b200/models.py: the feature converter (models.py:24-118) and the predict_batch wrapper
(models.py:121-152), on the CPU with a stand-in module."""
import numpy as np
import pytest

from mt3_b200 import models


def test_feature_converter_pads_trims_and_shifts():
    conv = models.ContinuousInputsEncDecFeatureConverter(pack=False)
    lengths = {'inputs': 6, 'targets': 5}
    short = {'inputs': np.arange(8, dtype=np.float32).reshape(4, 2) + 1, 'targets': np.array([7, 9, 1], np.int32)}
    out = conv.convert_example(short, lengths)
    assert out['encoder_input_tokens'].shape == (6, 2) and out['encoder_input_tokens'].dtype == np.float32
    np.testing.assert_array_equal(out['encoder_input_tokens'][:4], short['inputs'])
    assert (out['encoder_input_tokens'][4:] == 0.0).all()              # padding is 0.0 in log-mel space, not log(1e-5)
    np.testing.assert_array_equal(out['decoder_target_tokens'], [7, 9, 1, 0, 0])
    np.testing.assert_array_equal(out['decoder_input_tokens'], [0, 7, 9, 1, 0])      # seqio.autoregressive_inputs: shift right behind 0
    np.testing.assert_array_equal(out['decoder_loss_weights'], [1, 1, 1, 0, 0])      # seqio.non_padding_position
    long = {'inputs': np.ones((9, 2), np.float32), 'targets': np.arange(1, 9, dtype=np.int32)}
    out = conv.convert_example(long, lengths)
    assert out['encoder_input_tokens'].shape == (6, 2)
    np.testing.assert_array_equal(out['decoder_target_tokens'], [1, 2, 3, 4, 5])
    np.testing.assert_array_equal(out['decoder_input_tokens'], [0, 1, 2, 3, 4])
    # the inference datasets carry empty dummy targets (preprocessors.add_dummy_targets)
    out = conv.convert_example({'inputs': np.ones((6, 2), np.float32), 'targets': np.zeros((0,), np.int32)}, lengths)
    np.testing.assert_array_equal(out['decoder_input_tokens'], [0, 0, 0, 0, 0])
    assert out['decoder_loss_weights'].sum() == 0
    assert [o['encoder_input_tokens'].shape for o in conv([short, long], lengths)] == [(6, 2), (6, 2)]
    assert conv.get_model_feature_lengths({'inputs': 256, 'targets': 1024}) == {
        'encoder_input_tokens': 256, 'decoder_target_tokens': 1024, 'decoder_input_tokens': 1024, 'decoder_loss_weights': 1024}
    with pytest.raises(ValueError):
        models.ContinuousInputsEncDecFeatureConverter(pack=True)
    with pytest.raises(ValueError):
        conv.convert_example({'inputs': np.ones((6,), np.float32)}, {'inputs': 6, 'targets': 5})
    with pytest.raises(ValueError):
        conv.convert_example({'targets': np.ones((3,), np.int32)}, {'inputs': 6, 'targets': 5})


class _FakeModule:
    def __init__(self):
        self.calls = []

    def generate(self, x, stop_at_eos=True, decode='greedy'):
        self.calls.append((tuple(x.shape), stop_at_eos, decode))
        return np.zeros((x.shape[0], 8), np.int32)


def test_model_wrapper_predict_batch_contract():
    mod = _FakeModule()
    m = models.ContinuousInputsEncoderDecoderModel(mod, input_depth=4)                # decode_fn default = beam_search (models.py:127)
    batch = {'encoder_input_tokens': np.zeros((3, 5, 4), np.float32), 'decoder_input_tokens': np.zeros((3, 8), np.int32)}
    toks = m.predict_batch(None, batch)
    assert toks.shape == (3, 8) and mod.calls[-1] == ((3, 5, 4), True, 'beam1')
    toks, aux = m.predict_batch_with_aux(None, batch, decoder_params={'decode_rng': None})   # the notebook's call (:266-268)
    assert aux == {} and toks.shape == (3, 8)
    models.ContinuousInputsEncoderDecoderModel(mod, input_depth=4, decode_fn='greedy').predict_batch(None, batch)
    assert mod.calls[-1][2] == 'greedy'
    assert m.get_initial_variables(None, {'encoder_input_tokens': (8, 256), 'decoder_input_tokens': (8, 1024)}) == {
        'encoder_input_tokens': (8, 256, 4), 'decoder_input_tokens': (8, 1024)}
    assert m.get_initial_variables(None, {'encoder_input_tokens': (8, 256, 4)})['encoder_input_tokens'] == (8, 256, 4)
    with pytest.raises(AssertionError):
        m.get_initial_variables(None, {'encoder_input_tokens': (8, 256, 5)})
    with pytest.raises(ValueError):
        m.predict_batch(None, {'encoder_input_tokens': np.zeros((3, 5, 7), np.float32)})
    with pytest.raises(ValueError):
        m.predict_batch_with_aux(None, batch, num_decodes=4)
    with pytest.raises(ValueError):
        m.predict_batch_with_aux(None, batch, decoder_params={'decode_rng': 0})
    with pytest.raises(ValueError):
        models.ContinuousInputsEncoderDecoderModel(mod, decode_fn='sample')
    with pytest.raises(ValueError):
        models.ContinuousInputsEncoderDecoderModel(mod, label_smoothing=0.1)
