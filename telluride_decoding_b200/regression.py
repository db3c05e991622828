"""Regularisation sweep x leave-one-trial-out jackknife (telluride_decoding/regression.py).

The reference runs one complete fit per (lambda, held-out file): every job re-reads the data
and recomputes the whole Gram (regression.py:381-409 -> jackknife_one_model :197-214 ->
decoding.train_and_test).  Here the data is swept ONCE: each trial contributes its own
sufficient statistics, a fold's training statistics are total - held-out trial - the few rows
`drop_remainder` removes from that fold's stream (brain_data.py:369-370; a row-range Gram), and
every lambda of the sweep is solved from the same statistics.  Results are identical in
structure: `jackknife_over_regularizations` returns {lambda: (mean, std)} over files
(regression.py:416-420, population std).
"""

import collections
import re

import numpy as np
import torch

from telluride_decoding_b200 import brain_model
from telluride_decoding_b200 import cca
from telluride_decoding_b200 import decoding
from telluride_decoding_b200 import ops


def parse_regularization_values(mode_string):
  """Returns list of regularization values for different modes (regression.py:264-282)."""
  if isinstance(mode_string, float):
    return [mode_string,]
  if not isinstance(mode_string, str):
    raise TypeError('Parse_regularization_values needs a comma-separated' +
                    ' string, not a %s' % mode_string)
  mode_string = mode_string.lower()
  if mode_string == 'normal':
    return np.power(10, np.arange(-6.0, 0.5, 1))
  elif mode_string == 'test':
    return np.power(10, np.arange(-6.0, -5, 1))
  else:
    try:
      return np.array([float(tok) for tok in mode_string.split(',')], dtype=np.float32)
    except:   # pylint: disable=bare-except
      raise Exception('Could not parse regularization values: Want '
                      'comma separated list of floats, not %s' % mode_string)


def calculate_stats(run_results, axis=(1,)):
  """Mean and (population) std across held-out files (regression.py:245-261)."""
  return np.mean(run_results, axis=axis), np.std(run_results, axis=axis)


def get_brain_model(test_dataset, my_flags):
  if not isinstance(my_flags, decoding.DecodingOptions):
    raise TypeError('Get_brain_model needs a DecodingOptions object, ' +
                    'not %s.' % type(my_flags))
  return decoding.create_brain_model(my_flags, test_dataset)


def _single_file_dataset(brain_data_obj, name):
  saved = (brain_data_obj.train_file_pattern, brain_data_obj.validate_file_pattern,
           brain_data_obj.test_file_pattern)
  pattern = '^%s$' % re.escape(name)
  brain_data_obj.set_file_patterns('', pattern, pattern)
  try:
    return brain_data_obj.create_dataset('test')
  finally:
    brain_data_obj.set_file_patterns(*saved)


def _all_rows_stats(dataset, is_cca):
  """Statistics over EVERY row of the dataset's segments (no drop_remainder)."""
  keep = list(dataset.seg_rows)
  x = dataset.stream('x')
  if is_cca:
    y = dataset.stream('x2')
    g = ops.lagged_gram([x, y], keep)
    tx, n1 = ops.lagged_thin(x, None, keep)
    ty, _ = ops.lagged_thin(y, None, keep)
    return cca.CcaStats(g.to(torch.float64), tx[:, 0], ty[:, 0], float(n1.item()), x.width)
  g = ops.lagged_gram([x], keep)
  t, v = ops.lagged_thin(x, dataset.stream('y'), keep)
  return brain_model.RegressionStats(g.to(torch.float64), t, v)


def _dropped_rows_stats(train_dataset, is_cca):
  """Statistics of the rows drop_remainder removes from this fold's stream, or None."""
  if train_dataset.repeat_count != 1:
    raise ValueError('The single-pass jackknife expects repeat_count == 1.')
  kept = train_dataset.num_batches * train_dataset.final_batch_size
  if kept == train_dataset.total_rows:
    return None
  first = train_dataset.seg_keep_for_prefix(kept)
  last = list(train_dataset.seg_rows)
  x = train_dataset.stream('x')
  if is_cca:
    y = train_dataset.stream('x2')
    g = ops.lagged_gram_range([x, y], first, last)
    tx, n1 = ops.lagged_thin(x, None, last, seg_first=first)
    ty, _ = ops.lagged_thin(y, None, last, seg_first=first)
    return cca.CcaStats(g.to(torch.float64), tx[:, 0], ty[:, 0], float(n1.item()), x.width)
  g = ops.lagged_gram_range([x], first, last)
  t, v = ops.lagged_thin(x, train_dataset.stream('y'), last, seg_first=first)
  return brain_model.RegressionStats(g.to(torch.float64), t, v)


def jackknife_sweep(test_brain_data, my_flags, regularization_list, test_files=None,
                    solve_method=2):
  """[n_lambda, n_files] matrix of test metrics: one sweep over the data for all folds and
  lambdas.  Fold f trains on `allbut` {f} and tests on f, files in sorted order
  (regression.py:197)."""
  if not isinstance(my_flags, decoding.DecodingOptions):
    raise TypeError('jackknife_sweep needs a DecodingOptions object, not %s.' % type(my_flags))
  is_cca = my_flags.dnn_regressor == 'cca'
  files = sorted(test_files if test_files else test_brain_data.all_files())
  all_names = sorted(test_brain_data.all_files())
  lambdas = [float(l) for l in regularization_list]
  per_file = {}
  for name in all_names:                                   # the one pass over the data
    per_file[name] = _all_rows_stats(_single_file_dataset(test_brain_data, name), is_cca)
  total = None
  for name in all_names:
    s = per_file[name]
    total = s if total is None else (total + s if not is_cca else
                                     cca.CcaStats(total.gram + s.gram, total.sum_x + s.sum_x,
                                                  total.sum_y + s.sum_y, total.frames + s.frames,
                                                  s.d1))
  results = np.zeros((len(lambdas), len(files)), dtype=float)
  saved = (test_brain_data.train_file_pattern, test_brain_data.validate_file_pattern,
           test_brain_data.test_file_pattern)
  try:
    for j, name in enumerate(files):
      test_brain_data.set_file_patterns(my_flags.train_file_pattern or 'allbut', name, name)
      train_ds = test_brain_data.create_dataset('train')
      test_ds = test_brain_data.create_dataset('test')
      # The fold's training statistics = total - every file the train pattern leaves out (the
      # held-out file under `allbut`, more under allbut_NN or an explicit regexp; an explicit pattern
      # that does NOT exclude the held-out file keeps it, like the reference would).
      used = set(n for n, _, _ in train_ds.files)
      fold = total
      for other in all_names:
        if other not in used:
          fold = fold - per_file[other]
      dropped = _dropped_rows_stats(train_ds, is_cca)
      if dropped is not None:
        fold = fold - dropped
      model = decoding.create_brain_model(my_flags, test_ds)
      if is_cca:
        for i, lamb in enumerate(lambdas):
          model.fit_from_stats(fold, regularization_lambda=lamb)
          results[i, j] = model.evaluate(test_ds)[my_flags.test_metric]
      else:
        sols = brain_model.solve_regression(fold, lambdas, method=solve_method)
        for i, (w, b) in enumerate(sols):
          model.set_weights([w, b])
          results[i, j] = model.evaluate(test_ds)[my_flags.test_metric]
  finally:
    test_brain_data.set_file_patterns(*saved)
  return results


def jackknife_one_model(test_brain_data, test_brain_model, model_dir, my_flags,
                        max_test_count=-1, test_name='telluride4', trial_number=0,
                        summary_file=None, test_file=None):
  """One regularisation value, every held-out file (regression.py:151-242).  Returns the list
  of test metrics, one per file in sorted order."""
  del test_brain_model, model_dir, test_name, trial_number
  if not isinstance(my_flags, decoding.DecodingOptions):
    raise TypeError('Jackknife_one_model needs a DecodingOptions object, ' +
                    'not %s.' % type(my_flags))
  files = [test_file] if test_file else test_brain_data.all_files(max_test_count)
  res = jackknife_sweep(test_brain_data, my_flags, [my_flags.regularization_lambda], files)
  all_cor = list(res[0])
  if summary_file is not None and hasattr(summary_file, 'write'):
    summary_file.write('Jackknife test result regularization lambda={}, mean correlation={}, '
                       'std={}, test count={}\n'.format(my_flags.regularization_lambda,
                                                        np.mean(all_cor), np.std(all_cor),
                                                        len(all_cor)))
  return all_cor


class Regression(object):
  """Runs a series of experiments over a dataset (regression.py:285-443)."""

  def __init__(self, my_flags=None, brain_data_object=None):
    if my_flags and not isinstance(my_flags, decoding.DecodingOptions):
      raise TypeError('Regression init needs a DecodingOptions object, ' +
                      'not %s.' % type(my_flags))
    self.my_flags = my_flags or decoding.DecodingOptions()
    self.brain_data_object = brain_data_object
    self.test_name = 'Regression Test Object'

  @property
  def model_type(self):
    return 'Undefined'

  def preset_flags(self):
    """Generic presets (regression.py:422-443)."""
    self.my_flags.batch_norm = True
    self.my_flags.batch_size = 100
    self.my_flags.epoch_count = 1
    self.my_flags.input_field = 'eeg'
    self.my_flags.loss = 'mse'
    self.my_flags.output_field = 'intensity'
    self.my_flags.shuffle_buffer_size = 100
    if not self.my_flags.train_file_pattern:
      self.my_flags.train_file_pattern = 'allbut'
    return 'Generic'

  def jackknife_over_regularizations(self, my_flags, regularization_list, test_file=None,
                                     summary_base_dir='/tmp', model_base_dir=None,
                                     max_test_count=-1, brain_data_object=None):
    """{lambda: (mean, std)} over held-out files (regression.py:326-420)."""
    del summary_base_dir, model_base_dir
    data = brain_data_object or self.brain_data_object
    if data is None:
      raise ValueError('A BrainData object (TrialBrainData with the trials loaded) is required.')
    files = [test_file] if test_file else data.all_files(max_test_count)
    my_flags.train_file_pattern = my_flags.train_file_pattern or 'allbut'
    all_runs_results = jackknife_sweep(data, my_flags, regularization_list, files)
    self.all_runs_results = all_runs_results
    test_mean, test_std = calculate_stats(all_runs_results)
    mean_std_dict = collections.OrderedDict()
    for i in range(len(regularization_list)):
      mean_std_dict[regularization_list[i]] = (test_mean[i], test_std[i])
    return mean_std_dict


class RegressionLinear(Regression):
  """Linear presets (regression.py:446-460)."""

  @property
  def model_type(self):
    return 'linear'

  def preset_flags(self):
    super(RegressionLinear, self).preset_flags()
    self.my_flags.dnn_regressor = 'linear'
    if not self.my_flags.post_context:
      self.my_flags.post_context = 20
    self.my_flags.input2_pre_context = 0
    self.my_flags.input2_post_context = 0
    self.my_flags.input2_field = None
    self.my_flags.test_metric = 'pearson_correlation_first'
    self.my_flags.shuffle_buffer_size = 0
    return 'linear'


class RegressionCCA(Regression):
  """CCA presets (regression.py:463-482)."""

  @property
  def model_type(self):
    return 'cca'

  def preset_flags(self):
    super(RegressionCCA, self).preset_flags()
    self.my_flags.dnn_regressor = 'cca'
    if not self.my_flags.post_context:
      self.my_flags.post_context = 21
    if not self.my_flags.input2_pre_context:
      self.my_flags.input2_pre_context = 15
    if not self.my_flags.input2_post_context:
      self.my_flags.input2_post_context = 15
    self.my_flags.input2_field = 'intensity'
    self.my_flags.output_field = 'eeg'
    self.my_flags.test_metric = 'cca_pearson_correlation_first'
    self.my_flags.shuffle_buffer_size = 0
    self.my_flags.cca_dimensions = 5
    return 'cca'
