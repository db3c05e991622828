"""Attended-speaker decision rules (telluride_decoding/attention_decoder.py:116-173, :453-485).

Winner-take-all, the stepped (hysteresis) decoder and the state-space decoder.  The per-window
comparison of the two speakers' correlation means is `ops.wta_decide` on the GPU when whole
experiments are decoded at once; these classes keep the reference's streaming
`attention(r1, r2)` interface.  The state-space decoder (:176-451) is a serial scalar recursion
over a window of k_b + k_f + 1 correlation pairs (EM around a Kalman filter with Newton steps):
latency bound, a few hundred flops per decision, so it stays on the host in float64 like the
reference.  The plotting helper is out of scope.
"""

import numpy as np


class AttentionDecoder(object):
  """Winner takes all: is model 1 a better fit than model 2?  Strict greater-than."""

  def attention(self, r1, r2):
    return np.mean(r1) > np.mean(r2), 0, 0

  def tune(self, r1, r2):
    """An optional training step for tuning parameters."""
    del r1, r2


class StepAttentionDecoder(AttentionDecoder):
  """State starts at 0.5, moves +-0.1 per decision, clamped to [0.1, 0.9]; output state > 0.5."""

  def __init__(self):
    self.state = 0.5

  def attention(self, r1, r2):
    if np.mean(r1) > np.mean(r2):
      self.state = min(0.9, self.state + 0.1)
    else:
      self.state = max(0.1, self.state - 0.1)
    return self.state > 0.5, 0, 0


def step_decisions(wta):
  """The stepped decoder's outputs for a whole sequence of winner-take-all comparisons."""
  dec = StepAttentionDecoder()
  return np.array([dec.attention(1.0 if w else 0.0, 0.5)[0] for w in wta], dtype=bool)


def _log_normal_density(r, rho, mu):
  """Log-normal density up to the constant 1/sqrt(2 pi), precision rho, log-mean mu."""
  return (1.0 / r) * np.sqrt(rho) * np.exp(-0.5 * rho * (np.log(r) - mu) ** 2)


def _log_normal_fit(values):
  """(precision, log-mean) of a log-normal from the sample mean and (population) variance."""
  count = values.shape[0]
  mean = np.sum(values) / count
  variance = np.sum((values - mean) ** 2) / count
  precision = 1 / np.log(variance / mean ** 2 + 1)
  return precision, np.log(mean) - 0.5 / precision


class StateSpaceAttentionDecoder(AttentionDecoder):
  """Attention probability from a stream of correlation pairs (attention_decoder.py:176-451).

  Model: |r| of the attended (unattended) speaker is log-normal with parameters (mu_d[0], rho_d[0])
  ((mu_d[1], rho_d[1])); the attention logit z_k is a random walk with variance eta_k under an
  inverse-gamma prior.  Every call, once k_w = forward_lag + backward_lag + 1 pairs have been seen,
  runs `outer_iter` EM rounds over the last k_w pairs: posterior attention eps_k (E), log-normal
  parameters (M), then `inner_iter` passes of {Kalman filter with `newton_iter` Newton steps per
  sample, smoother, variance update}; it returns (p, p_low, p_high) for the sample `forward_lag`
  frames back, c0 = 1.96 standard deviations apart.  The arithmetic, including the filter state
  carried from call to call and the smoother sweeping the window front to back (:425-432), follows
  the reference operation by operation so that the outputs agree to rounding.
  """

  def __init__(self, outer_iter, inner_iter, newton_iter, fs_corr, forward_lag=0, backward_lag=13,
               offset=0.0):
    self._offset = offset
    self.outer_iter, self.inner_iter, self.newton_iter = outer_iter, inner_iter, newton_iter
    self.fs_corr = fs_corr
    self.forward_lag, self.backward_lag = forward_lag, backward_lag
    self.k_f, self.k_b = forward_lag, backward_lag
    self.k_w = self.k_f + self.k_b + 1
    self.c0 = 1.96                                  # 95 % confidence band
    self.mean_p, self.var_p = 0.2, 5                # inverse-gamma prior of the walk variances
    self.a_0 = 2 + self.mean_p ** 2 / self.var_p
    self.b_0 = self.mean_p * (self.a_0 - 1)
    self.calls = 0
    self.r1, self.r2 = [], []
    self.z_smoothed = [0.0] * self.k_w
    self.eta_smoothed = [0.3] * self.k_w
    self.z_dyn = [0.0] * self.k_w
    self.eta_dyn = [0.0] * self.k_w
    self.lambda_state = 1.0
    n = self.k_w + 1
    self.z_k_k, self.sig_k_k = np.zeros((n,)), np.zeros((n,))           # filtered
    self.z_k_k_1, self.sig_k_k_1 = np.zeros((n,)), np.zeros((n,))       # one-step predictions
    self.z_k_k_cap, self.sig_k_k_cap = np.zeros((n,)), np.zeros((n,))   # smoothed
    self.sm = np.zeros((self.k_w,))
    # Tuned prior hyperparameters of the attended / unattended log-normal distributions (:268-274).
    self.alpha_0 = [6.4113e+02, 4.0434e+03]
    self.beta_0 = [3.7581e+02, 6.2791e+03]
    self.mu_0 = [-0.3994, -1.5103]
    self.rho_d = [1.7060, 0.64395]
    self.mu_d = [-0.3994, -1.5103]

  def tune(self, r1, r2):
    return self.tune_log_normal_priors(r1, r2)

  def tune_log_normal_priors(self, r1, r2):
    """Moment estimates of both log-normals from training correlations; r1 MUST be the attended
    speaker (attention_decoder.py:281-327)."""
    rho_a, mu_a = _log_normal_fit(np.absolute(np.asarray(r1) + self._offset))
    rho_u, mu_u = _log_normal_fit(np.absolute(np.asarray(r2) + self._offset))
    self.rho_d = [rho_a, rho_u]
    self.mu_d = [mu_a, mu_u]
    self.mu_0 = [mu_a, mu_u]
    self.alpha_0 = [6.4113e+02, 4.0434e+03]
    self.beta_0 = [3.7581e+02, 6.2791e+03]

  # ---- one EM round -------------------------------------------------------------------------
  def _posterior_attention(self, r1, r2, z):
    """E step: probability that speaker 1 is the attended one, per sample (:362-378)."""
    p_11 = _log_normal_density(r1, self.rho_d[0], self.mu_d[0])
    p_12 = _log_normal_density(r1, self.rho_d[1], self.mu_d[1])
    p_21 = _log_normal_density(r2, self.rho_d[1], self.mu_d[1])
    p_22 = _log_normal_density(r2, self.rho_d[0], self.mu_d[0])
    p = 1.0 / (1.0 + np.exp(-z))
    return (p * p_11 * p_21) / (p * p_11 * p_21 + (1.0 - p) * p_12 * p_22)

  def _update_log_normals(self, r1, r2, ep):
    """M step for (mu_d, rho_d) with their priors (:380-396)."""
    k_w = self.k_w
    self.mu_d[0] = (np.sum(ep * np.log(r1) + (1.0 - ep) * np.log(r2)) + k_w * self.mu_0[0]) / (2.0 * k_w)
    self.mu_d[1] = (np.sum(ep * np.log(r2) + (1.0 - ep) * np.log(r1)) + k_w * self.mu_0[1]) / (2.0 * k_w)
    self.rho_d[0] = (2.0 * k_w * self.alpha_0[0]) / (
        np.sum(ep * ((np.log(r1) - self.mu_d[0]) ** 2) + (1.0 - ep) * ((np.log(r2) - self.mu_d[0]) ** 2)) +
        k_w * (2.0 * self.beta_0[0] + (self.mu_d[0] - self.mu_0[0]) ** 2))
    self.rho_d[1] = (2.0 * k_w * self.alpha_0[1]) / (
        np.sum(ep * ((np.log(r2) - self.mu_d[1]) ** 2) + (1.0 - ep) * ((np.log(r1) - self.mu_d[1]) ** 2)) +
        k_w * (2.0 * self.beta_0[1] + (self.mu_d[1] - self.mu_0[1]) ** 2))

  def _filter(self, ep, eta):
    """Forward pass: prediction, Newton iterations on the Bernoulli-logit posterior mode, posterior
    variance (:400-419).  z_k_k[k] starts from its value of the previous pass."""
    lam = self.lambda_state
    for k in range(1, self.k_w + 1):
      self.z_k_k_1[k] = lam * self.z_k_k[k - 1]
      self.sig_k_k_1[k] = lam ** 2 * self.sig_k_k[k - 1] + eta[k - 1]
      z_pred, s_pred = self.z_k_k_1[k], self.sig_k_k_1[k]
      for _ in range(self.newton_iter):
        z = self.z_k_k[k]
        ez = np.exp(z)
        self.z_k_k[k] = z - (z - z_pred - s_pred * (ep[k - 1] - ez / (1 + ez))) / (
            1 + s_pred * ez / ((1 + ez) ** 2))
      ez = np.exp(self.z_k_k[k])
      self.sig_k_k[k] = 1.0 / (1.0 / s_pred + ez / ((1 + ez) ** 2))

  def _smooth(self):
    """Smoother gains and smoothed moments, swept from the front of the window (:421-435), then the
    walk variances under the inverse-gamma prior (:437-439)."""
    k_w = self.k_w
    self.z_k_k_cap[k_w] = self.z_k_k[k_w]
    self.sig_k_k_cap[k_w] = self.sig_k_k[k_w]
    for k in range(k_w):
      gain = self.sig_k_k[k] * self.lambda_state / self.sig_k_k_1[k + 1]
      self.sm[k] = gain
      self.z_k_k_cap[k] = self.z_k_k[k] + gain * (self.z_k_k_cap[k + 1] - self.z_k_k_1[k + 1])
      self.sig_k_k_cap[k] = self.sig_k_k[k] + gain ** 2 * (self.sig_k_k_cap[k + 1] - self.sig_k_k_1[k + 1])
    self.z_k_k[0] = self.z_k_k_cap[0]
    self.sig_k_k[0] = self.sig_k_k_cap[0]
    return ((self.z_k_k_cap[1:] - self.z_k_k_cap[:-1]) ** 2 + self.sig_k_k_cap[1:] +
            self.sig_k_k_cap[:-1] - 2.0 * self.sig_k_k_cap[1:] * self.sm + 2 * self.b_0) / (
                1 + 2 * (self.a_0 + 1))

  def attention(self, r1, r2):
    """(mean, lower bound, upper bound) of the attention probability after two new correlations;
    (0.5, 0.5, 0.5) until k_w pairs have arrived (:329-451)."""
    self.calls += 1
    self.r1.append(np.abs(r1 + self._offset))
    self.r2.append(np.abs(r2 + self._offset))
    if self.calls < self.k_w:
      return (0.5, 0.5, 0.5)
    k_w = self.k_w
    w1, w2 = np.array(self.r1[-k_w:]), np.array(self.r2[-k_w:])
    z = np.array(self.z_smoothed[-k_w:])
    eta = np.array(self.eta_smoothed[-k_w:])
    for _ in range(self.outer_iter):
      ep = self._posterior_attention(w1, w2, z)
      self._update_log_normals(w1, w2, ep)
      for _ in range(self.inner_iter):
        self._filter(ep, eta)
        eta = self._smooth()
      z = self.z_k_k_cap[1:]
    self.z_smoothed += list(self.z_k_k_cap[1:])
    self.eta_smoothed += list(eta)
    self.z_k_k[0] = self.z_k_k_cap[1]
    self.z_dyn.append(self.z_smoothed[-1 - self.k_f])
    self.eta_dyn.append(self.eta_smoothed[-1 - self.k_f])
    z_now, spread = self.z_dyn[-1], self.c0 * np.sqrt(self.eta_dyn[-1])
    return (1.0 / (1 + np.exp(-z_now)), 1.0 / (1 + np.exp(-z_now - spread)),
            1.0 / (1 + np.exp(-z_now + spread)))


def create_attention_decoder(type_name, window_step=100, frame_rate=100.0, ssd_offset=0.0):
  """'wta', 'stepped' / 'step' or 'ssd' (attention_decoder.py:453-485)."""
  if type_name == 'wta':
    return AttentionDecoder()
  elif type_name == 'stepped' or type_name == 'step':
    return StepAttentionDecoder()
  elif type_name == 'ssd':
    fs_corr = window_step * float(frame_rate) / 2.0
    return StateSpaceAttentionDecoder(20, 1, 10, fs_corr, offset=ssd_offset)
  raise ValueError('Unknown type (%s) requested from create_attention_decoder' % type_name)
