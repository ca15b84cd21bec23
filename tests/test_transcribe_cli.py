"""`python -m mt3_b200.transcribe`: WAV in, MIDI out (the notebook's upload -> transcribe -> download cells as one command)."""
import json
import struct

import numpy as np
import pytest

from mt3_b200 import audio_io, transcribe


def _wav(path, seconds=5.0, rate=44100, channels=2):
    t = np.arange(int(seconds * rate)) / rate
    x = 0.4 * np.sin(2 * np.pi * 261.63 * t) + 0.3 * np.sin(2 * np.pi * 392.0 * t)
    pcm = (np.stack([x] * channels, axis=1) * 32767).round().astype('<i2').tobytes()
    hdr = b'RIFF' + struct.pack('<I', 36 + len(pcm)) + b'WAVE' + b'fmt ' + struct.pack('<IHHIIHH', 16, 1, channels, rate,
                                                                                      rate * channels * 2, channels * 2, 16)
    path.write_bytes(hdr + b'data' + struct.pack('<I', len(pcm)) + pcm)


def test_cli_argument_and_input_errors(tmp_path, capsys):
    with pytest.raises(SystemExit):
        transcribe.main([])                                           # audio / midi / --checkpoint are required
    with pytest.raises(SystemExit):
        transcribe.main(['a.wav', 'b.mid', '--checkpoint', 'synthetic', '--model', 'nope'])
    assert transcribe.main([str(tmp_path / 'missing.wav'), str(tmp_path / 'o.mid'), '--checkpoint', 'synthetic']) == 2
    bad = tmp_path / 'bad.wav'
    bad.write_bytes(b'ID3\x00not a wave file')
    assert transcribe.main([str(bad), str(tmp_path / 'o.mid'), '--checkpoint', 'synthetic']) == 2
    assert 'cannot read' in capsys.readouterr().err


def test_cli_fails_loudly_without_a_gpu(tmp_path, capsys):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    _wav(tmp_path / 'in.wav', seconds=0.5)
    assert transcribe.main([str(tmp_path / 'in.wav'), str(tmp_path / 'o.mid'), '--checkpoint', 'synthetic']) == 3
    assert 'no CPU fallback' in capsys.readouterr().err
    assert not (tmp_path / 'o.mid').exists()


@pytest.mark.gpu
def test_cli_wav_to_midi_equals_the_api_calls(tmp_path, capsys):
    """5 s of 44.1 kHz stereo -> resample -> 3 mt3 segments -> tokens -> stitched notes -> SMF: the file the command writes is
    byte-identical to the one built by calling the pieces by hand (InferenceModel.__call__ + note_sequence_to_midi_bytes)."""
    from mt3_b200 import inference, note_decoding
    _wav(tmp_path / 'in.wav')
    rc = transcribe.main([str(tmp_path / 'in.wav'), str(tmp_path / 'out.mid'), '--checkpoint', 'synthetic:3', '--batch-size', '4',
                          '--jsonl', str(tmp_path / 'notes.jsonl')])
    assert rc == 0
    summary = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert summary['audio_seconds'] == pytest.approx(5.0, abs=1e-3)
    audio = audio_io.load_audio(str(tmp_path / 'in.wav'), 16000)
    assert audio.shape == (80000,)
    model = inference.InferenceModel('synthetic:3', 'mt3', batch_size=4)
    ns = model(audio)
    data = (tmp_path / 'out.mid').read_bytes()
    assert data[:4] == b'MThd' and data == note_decoding.note_sequence_to_midi_bytes(ns)
    rec = json.loads((tmp_path / 'notes.jsonl').read_text())
    assert rec['id'] == 'in.wav' and len(rec['est_notes']) == len(ns.notes) == summary['numNotes'] + summary['numDrumNotes']
