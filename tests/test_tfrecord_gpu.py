"""TFExampleData end to end on the GPU: the reference's MEG fixture bytes -> native parser -> lagged
kernels -> ridge, against the unmodified reference function (tests/golden/meg_subj01_head.npz), and
several files against the in-memory TrialBrainData (same interleave / context semantics)."""
import os
import shutil

import numpy as np
import pytest

from oracle import ref_numpy as ora
from telluride_decoding_b200 import brain_data, brain_model, tfrecord
from tests import gpu_util

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
HEAD = os.path.join(GOLDEN, 'meg_subj01_head.tfrecords')


def test_meg_fixture_ridge_matches_reference(tmp_path):
  g = np.load(os.path.join(GOLDEN, 'meg_subj01_head.npz'))
  shutil.copy(HEAD, str(tmp_path / 'subj01_head.tfrecords'))
  bd = brain_data.create_brain_dataset('tfrecord', 'mel_spectrogram', 'envelope', 100.0,
                                       pre_context=int(g['pre']), post_context=int(g['post']),
                                       final_batch_size=int(g['batch']), shuffle_buffer_size=0,
                                       data_dir=str(tmp_path))
  ds = bd.create_dataset('train')
  assert bd.input_fields_width() == 64 * 4 and bd.output_field_width() == 1
  w, b, _, _, _ = brain_model.calculate_linear_regressor_parameters_from_dataset(ds, lamb=0.1)
  assert gpu_util.rel_fro(w, g['ridge_w']) < 1e-4
  assert np.allclose(b, g['ridge_b'], rtol=1e-4, atol=1e-4 * np.abs(g['ridge_b']).max())
  # the batches are the reference's lagged rows, bit for bit
  zeros = np.zeros((300, 1), np.float32)
  rows = ora.file_rows(g['col_mel_spectrogram'], g['col_mel_spectrogram'][:, 0:1], g['col_envelope'],
                       zeros, int(g['pre']), int(g['post']))
  want = ora.batches_from_files([rows], int(g['batch']))
  got = list(ds)
  assert len(got) == len(want) == 6
  for (gd, gy), (od, oy) in zip(got, want):
    np.testing.assert_array_equal(gd['input_1'].cpu().numpy(), od['input_1'])
    np.testing.assert_array_equal(gy.cpu().numpy(), oy)


def test_several_files_equal_trials_in_memory(tmp_path):
  rng = np.random.default_rng(11)
  trials = {}
  for i, n in enumerate((310, 275, 420)):
    eeg = rng.standard_normal((n, 6)).astype(np.float32)
    spec = rng.standard_normal((n, 3)).astype(np.float32)
    env = (eeg[:, :1] * 0.5 + 0.1 * rng.standard_normal((n, 1))).astype(np.float32)
    att = np.full((n, 1), float(i % 2), np.float32)
    name = 'trial_%02d.tfrecords' % i
    trials[name] = (eeg, spec, env, att)
    tfrecord.write_tfrecord(str(tmp_path / name), {'eeg': eeg, 'spectrogram': spec, 'envelope': env,
                                                  'attended': att})
  kwargs = dict(pre_context=2, post_context=1, in2_fields='spectrogram', in2_pre_context=0,
                in2_post_context=2, attended_field='attended', final_batch_size=100,
                shuffle_buffer_size=0)
  disk = brain_data.TFExampleData('eeg', 'envelope', 64.0, data_dir=str(tmp_path), **kwargs)
  mem = brain_data.TrialBrainData('eeg', 'envelope', 64.0, **kwargs)
  for name, (eeg, spec, env, att) in trials.items():
    mem.add_trial(os.path.join(str(tmp_path), name), eeg, env, spec, att)
  mem._cached_file_names = list(disk.all_files())        # same (shuffled) file order
  a, b = list(disk.create_dataset('train')), list(mem.create_dataset('train'))
  assert len(a) == len(b) == (310 + 275 + 420) // 100
  for (da, ya), (db, yb) in zip(a, b):
    for k in ('input_1', 'input_2', 'attended_speaker'):
      np.testing.assert_array_equal(da[k].cpu().numpy(), db[k].cpu().numpy())
    np.testing.assert_array_equal(ya.cpu().numpy(), yb.cpu().numpy())
  # 'ones' output (brain_data.py:778-780) and the CCA model on files
  from telluride_decoding_b200 import cca
  disk2 = brain_data.TFExampleData('eeg', 'ones', 64.0, data_dir=str(tmp_path), **kwargs)
  ds2 = disk2.create_dataset('train')
  (_, y_first), = ds2.take(1)
  assert float(y_first.min()) == float(y_first.max()) == 1.0
  model = cca.BrainModelCCA(ds2, cca_dims=2, regularization_lambda=0.1)
  model.fit(ds2)
  assert model.evaluate(ds2)['cca_pearson_correlation_first'] > -1.0
