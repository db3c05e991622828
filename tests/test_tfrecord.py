"""The native TFRecord / tf.train.Example reader (host code of libtdgram.so) against the bytes of
the reference's own MEG fixture (test/brain_data_test.py:509-526) and a writer round trip."""
import os
import shutil

import numpy as np
import pytest

from telluride_decoding_b200 import brain_data
from telluride_decoding_b200 import tfrecord
from telluride_decoding_b200._lib import TdError

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
HEAD = os.path.join(GOLDEN, 'meg_subj01_head.tfrecords')


def test_reads_the_reference_fixture_bit_exactly():
  g = np.load(os.path.join(GOLDEN, 'meg_subj01_head.npz'))
  assert tfrecord.count_tfrecords(HEAD) == (300, False)
  table = tfrecord.feature_table(HEAD)
  # widths of test/brain_data_test.py:521-526
  assert dict((n, w) for n, w, _ in table) == {'meg': 148, 'mel_spectrogram': 64, 'envelope': 1,
                                                'phonemes': 38, 'phonetic_features': 19}
  assert all(kind == 2 for _, _, kind in table)
  shapes = brain_data.discover_feature_shapes(HEAD)
  assert shapes['meg'].shape == [148] and shapes['envelope'].shape == [1]
  names = ['meg', 'envelope', 'mel_spectrogram', 'phonemes', 'phonetic_features']
  got = tfrecord.read_features(HEAD, names, verify_crc=2, pin=False).numpy()
  want = np.concatenate([g['col_' + n] for n in names], axis=1)
  np.testing.assert_array_equal(got, want)
  # any order, any subset
  sub = tfrecord.read_features(HEAD, ['envelope', 'meg'], pin=False).numpy()
  np.testing.assert_array_equal(sub[:, 0:1], g['col_envelope'])
  np.testing.assert_array_equal(sub[:, 1:], g['col_meg'])


def test_errors(tmp_path):
  with pytest.raises(ValueError, match='Could not find all desired features'):
    tfrecord.read_features(HEAD, ['meg', 'no_such_field'], pin=False)
  with pytest.raises(TdError, match='does not have 100 values'):
    tfrecord.read_features(HEAD, ['meg'], widths=[100], pin=False)
  with pytest.raises(TypeError, match='must be a string'):
    brain_data.count_tfrecords(12)
  with pytest.raises(TypeError, match='input must be a string filename'):
    brain_data.discover_feature_shapes(None)
  raw = bytearray(open(HEAD, 'rb').read())
  raw[1212 * 7 + 500] ^= 0x40                        # inside the payload of record 7
  bad = str(tmp_path / 'damaged.tfrecords')
  open(bad, 'wb').write(raw)
  assert tfrecord.count_tfrecords(bad) == (7, True)
  assert tfrecord.count_tfrecords(bad, verify_crc=1) == (300, False)     # framing is intact
  open(bad, 'wb').write(raw[:1212 * 3 + 100])        # truncated in record 3
  assert tfrecord.count_tfrecords(bad) == (3, True)
  with pytest.raises(TdError, match='truncated record 3'):
    tfrecord.read_features(bad, ['meg'], widths=[148], pin=False)


def test_writer_round_trip(tmp_path):
  rng = np.random.default_rng(3)
  eeg = rng.standard_normal((77, 9)).astype(np.float32)
  env = rng.standard_normal((77, 1)).astype(np.float32)
  path = str(tmp_path / 'sub' / 'trial_a.tfrecords')
  tfrecord.write_tfrecord(path, {'eeg': eeg, 'envelope': env})
  assert tfrecord.count_tfrecords(path) == (77, False)
  got = tfrecord.read_features(path, ['eeg', 'envelope'], verify_crc=2, pin=False).numpy()
  np.testing.assert_array_equal(got[:, :9], eeg)
  np.testing.assert_array_equal(got[:, 9:], env)


def test_file_discovery_like_the_reference(tmp_path):
  """brain_data.py:648-689: *.tfrecords below data_dir that contain data_pattern, no '-bad-'."""
  for name in ('subj01_a.tfrecords', 'subj02_a.tfrecords', 'subj03-bad-.tfrecords', 'other.txt'):
    shutil.copy(HEAD, str(tmp_path / name))
  os.makedirs(str(tmp_path / 'deeper'))
  shutil.copy(HEAD, str(tmp_path / 'deeper' / 'subj04_a.tfrecords'))
  bd = brain_data.TFExampleData('meg', 'envelope', 100.0, data_dir=str(tmp_path), data_pattern='subj',
                                train_file_pattern='subj0[12]', test_file_pattern='subj04')
  found = sorted(os.path.basename(f) for f in bd.all_files())
  assert found == ['subj01_a.tfrecords', 'subj02_a.tfrecords', 'subj04_a.tfrecords']
  assert sorted(os.path.basename(f) for f in bd.filter_file_names('train')) == [
      'subj01_a.tfrecords', 'subj02_a.tfrecords']
  assert [os.path.basename(f) for f in bd.filter_file_names('test')] == ['subj04_a.tfrecords']
  assert bd.features['meg'].shape == [148]
  with pytest.raises(ValueError, match='Missing data_dir in TFExampleData initialization'):
    brain_data.TFExampleData('meg', 'envelope', 100.0)
  with pytest.raises(ValueError, match='Should not have an empty list of data files'):
    brain_data.TFExampleData('meg', 'envelope', 100.0, data_dir=str(tmp_path), data_pattern='zzz')
  assert isinstance(brain_data.create_brain_dataset('tfrecord', 'meg', 'envelope', 100.0,
                                                    data_dir=str(tmp_path)), brain_data.TFExampleData)
