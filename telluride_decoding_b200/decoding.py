"""Experiment options and the train-and-test driver (telluride_decoding/decoding.py).

`DecodingOptions` keeps the reference's field names and defaults (decoding.py:48-92) as a plain
class (no absl flags on the hot path); `create_brain_model` dispatches 'linear' and 'cca'
(decoding.py:259-311) -- the SGD-trained DNN / classifier models are outside the dense decoding
path; `train_and_test` is decoding.py:314-350.
"""

from telluride_decoding_b200 import brain_data
from telluride_decoding_b200 import brain_model
from telluride_decoding_b200 import cca


class DecodingOptions(object):
  """All the parameters of a decoding experiment (decoding.py:48-154)."""

  _DEFAULTS = dict(
      attended_field='attend', batch_norm=False, batch_size=512, cca_dimensions=10,
      check_file_pattern='', correlation_frames=100, correlation_reducer='lda', data='tfrecords',
      debug=False, dnn_regressor='fullyconnected', dropout=0.0, epoch_count=100, frame_rate=100.0,
      hidden_units='20-20', input2_field='', input2_post_context=0, input2_pre_context=0,
      input_offset=0, input_field='mel_spectrogram', learning_rate=0.05, loss='mse',
      min_context=0, output_field='envelope', post_context=0, pre_context=0,
      random_mixup_batch=False, regularization_lambda=0.1, saved_model_dir=None,
      shuffle_buffer_size=100000, summary_dir='/tmp/tf', tensorboard_dir=None,
      test_file_pattern='', test_metric='pearson_correlation_first', tfexample_dir=None,
      tfexample_pattern='', train_file_pattern='', validate_file_pattern='')

  def __init__(self, **overrides):
    for k, v in self._DEFAULTS.items():
      setattr(self, k, v)
    self.set_from_dict(overrides)

  def set_flags(self, all_flags=None):
    """Copies same-named attributes from a flags-like object (decoding.py:94-134)."""
    if all_flags is not None:
      for k in self._DEFAULTS:
        if hasattr(all_flags, k):
          setattr(self, k, getattr(all_flags, k))
    return self

  def experiment_parameters(self, delimiter=','):
    keys_and_values = ['%s=%s' % (k, getattr(self, k)) for k in sorted(self._DEFAULTS)]
    if delimiter:
      return delimiter.join(keys_and_values)
    return keys_and_values

  def set_from_dict(self, new_values):
    for k, v in new_values.items():
      setattr(self, k, v)
    return self


def create_brain_model(model_flags, input_dataset):
  """Creates the right kind of brain model (decoding.py:259-311)."""
  if not isinstance(model_flags, DecodingOptions):
    raise TypeError('Model_flags must be a DecodingOptions, not a %s' % type(model_flags))
  if not brain_model.is_dataset(input_dataset):
    raise TypeError('input_dataset must be a tf.data.Dataset, not %s' % type(input_dataset))
  if model_flags.dnn_regressor == 'linear':
    bm = brain_model.BrainModelLinearRegression(
        input_dataset, model_flags.regularization_lambda,
        tensorboard_dir=model_flags.tensorboard_dir)
  elif model_flags.dnn_regressor == 'cca':
    bm = cca.BrainModelCCA(
        input_dataset, cca_dims=model_flags.cca_dimensions,
        regularization_lambda=model_flags.regularization_lambda,
        tensorboard_dir=model_flags.tensorboard_dir)
  elif model_flags.dnn_regressor in ('fullyconnected', 'classifier'):
    raise TypeError('Model type %s is trained by SGD in the reference and is outside the '
                    'B200 dense decoding path (linear and cca are provided).' %
                    model_flags.dnn_regressor)
  else:
    raise TypeError('Unknown model type %s in create_brain_model.' % model_flags.dnn_regressor)
  bm.compile(learning_rate=model_flags.learning_rate)
  return bm


def train_and_test(my_flags, test_brain_data, test_brain_model, epochs=1):
  """Trains on the 'train' files and evaluates on the 'test' files (decoding.py:314-350)."""
  if not isinstance(test_brain_data, brain_data.BrainData):
    raise TypeError('test_brain_data must be a BrainData object, not a %s' % test_brain_data)
  if not isinstance(test_brain_model, brain_model.BrainModel):
    raise TypeError('Model in train_and_test must be a BrainModel object, '
                    'not %s' % test_brain_model)
  if not isinstance(my_flags, DecodingOptions):
    raise TypeError('Train_and_test needs a DecodingOptions object, not %s.' % type(my_flags))
  train_dataset = test_brain_data.create_dataset('train')
  train_results = test_brain_model.fit(train_dataset, epochs=epochs)
  test_dataset = test_brain_data.create_dataset('test')
  test_results = test_brain_model.evaluate(test_dataset)
  return train_results, test_results
