"""CPU restatement (numpy) of the two non-tree classifiers of the reference scripts — TEST INFRASTRUCTURE ONLY, like the rest
of oracle/ (only tests/ may import it).  PARITY UNPINNED at the MLlib boundary: Spark is absent (no JVM) and the reference holds
no vectors; the independent pins are scikit-learn's MultinomialNB (theta) and LogisticRegression(saga, elastic-net) (minimiser).

Follows [MLlib, recalled]: ml/classification/NaiveBayes.scala (trainWithLabelCheck: pi, theta), ml/classification/
LogisticRegression.scala (standardised features, unpenalised intercepts, regParam * (alpha |B|_1 + (1 - alpha)/2 |B|^2)),
called from kdd99.py:57-58,67 and cicids17.py:61-62,71.
"""
import numpy as np


def nb_fit(x, y, num_classes, smoothing=1.0):
    x = np.asarray(x, np.float64); y = np.asarray(y, np.int64)
    if (x < 0).any():
        raise ValueError("Naive Bayes requires nonnegative feature values")
    C, D = int(num_classes), x.shape[1]
    pi = np.full(C, -np.inf); theta = np.zeros((C, D))
    present = [c for c in range(C) if (y == c).any()]
    L, N = len(present), len(y)
    for c in present:
        rows = x[y == c]
        n_c, s = rows.shape[0], rows.sum(0)
        pi[c] = np.log(n_c + smoothing) - np.log(N + L * smoothing)
        theta[c] = np.log(s + smoothing) - np.log(s.sum() + D * smoothing)
    return pi, theta


def nb_predict(pi, theta, x):
    raw = np.asarray(x, np.float64) @ theta.T + pi
    m = raw.max(1, keepdims=True)
    e = np.exp(raw - m)
    return raw, e / e.sum(1, keepdims=True), raw.argmax(1).astype(np.float64)


def standardize(x):
    x = np.asarray(x, np.float64)
    std = x.std(0, ddof=1)
    inv = np.where(std > 0, 1.0 / np.where(std > 0, std, 1.0), 0.0)
    return x * inv, inv


def lr_objective(xs, y, B, b, reg_param, alpha):
    """B [C, D] on STANDARDISED features xs, b [C]: mean multinomial log-loss + elastic-net penalty on B."""
    z = xs @ B.T + b
    m = z.max(1, keepdims=True)
    lse = (m + np.log(np.exp(z - m).sum(1, keepdims=True)))[:, 0]
    loss = (lse - z[np.arange(len(y)), y]).mean()
    return loss + reg_param * (alpha * np.abs(B).sum() + 0.5 * (1.0 - alpha) * (B * B).sum())


def lr_minimise_ista(xs, y, num_classes, reg_param, alpha, iters=20000):
    """reference minimiser: plain proximal gradient with the 1/L step — slow, simple, independent of the product's OWL-QN."""
    n, D = xs.shape; C = int(num_classes)
    Y = np.eye(C)[y]
    B = np.zeros((C, D)); b = np.zeros(C)
    l1, l2 = reg_param * alpha, reg_param * (1.0 - alpha)
    Lc = 0.5 * (np.linalg.norm(xs, 2) ** 2 / n + 1.0) + l2
    for _ in range(iters):
        z = xs @ B.T + b
        z -= z.max(1, keepdims=True)
        P = np.exp(z); P /= P.sum(1, keepdims=True)
        R = (P - Y) / n
        gB, gb = R.T @ xs + l2 * B, R.sum(0)
        B = B - gB / Lc
        B = np.sign(B) * np.maximum(np.abs(B) - l1 / Lc, 0.0)
        b = b - gb / Lc
    return B, b - b.mean()
