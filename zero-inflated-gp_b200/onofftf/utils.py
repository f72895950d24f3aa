"""Host-side recipes around the Kronecker trainer (no GPU work): the reference scripts' cross-validation splits, inducing-point
initialisation, error summaries and the two-stage combiners of the comparison models.

    cv_splits              scripts/create_cvsplits.py:14-34   (5-fold KFold, shuffle, random_state 1234, time column / 1000)
    init_inducing          scripts/onoff.py:67-68             (k-means on the spatial columns, linspace over the time column)
    rmse / mae             scripts/onoff.py:471-481, scripts/zero_inflated.py:62-68   (predictions clipped at zero first)
    hurdle_select          scripts/hurdle.py:47-54            (rows the classifier calls "on": p > 0.5)
    zero_inflated_combine  scripts/zero_inflated.py:50-60     (classifier probability or hard cut times the regression mean)
"""
from __future__ import annotations

import numpy as np


def cv_splits(X, Y, n_splits=5, random_state=1234, time_column=2, time_scale=1000.0):
    """Yields (fold, dict(Xtrain, Ytrain, Xtest, Ytest)), fold = 1..n_splits, as scripts/create_cvsplits.py writes them.
    The time column is divided by `time_scale` before splitting (create_cvsplits.py:16)."""
    from sklearn.model_selection import KFold
    X = np.array(X, dtype=float, copy=True)
    Y = np.asarray(Y, dtype=float)
    if time_column is not None:
        X[:, time_column] = X[:, time_column] / time_scale
    kf = KFold(n_splits=n_splits, random_state=random_state, shuffle=True)
    for fold, (tr, te) in enumerate(kf.split(X), start=1):
        yield fold, {"Xtrain": X[tr], "Ytrain": Y[tr], "Xtest": X[te], "Ytest": Y[te]}


def init_inducing(Xtrain, num_inducing=(10, 100), space_columns=(0, 1), time_column=2, seed=None):
    """[Z_space (Ms x 2) from k-means of the spatial columns, Z_time (Mt x 1) equally spaced over the observed time range]."""
    from scipy.cluster.vq import kmeans
    Xtrain = np.asarray(Xtrain, dtype=float)
    Zs = kmeans(Xtrain[:, list(space_columns)], int(num_inducing[0]), **({"seed": seed} if seed is not None else {}))[0]
    t = Xtrain[:, time_column]
    Zt = np.expand_dims(np.linspace(t.min(), t.max(), int(num_inducing[1])), axis=1)
    return [Zs, Zt]


def rmse(predict, actual):
    predict = np.maximum(np.asarray(predict, dtype=float), 0.0)
    return float(np.sqrt(np.mean((np.asarray(actual, dtype=float) - predict) ** 2)))


def mae(predict, actual):
    predict = np.maximum(np.asarray(predict, dtype=float), 0.0)
    return float(np.mean(np.abs(np.asarray(actual, dtype=float) - predict)))


def hurdle_select(clf_prob, threshold=0.5):
    """row indices the classifier predicts as non-zero; the hurdle regression is trained and evaluated on these rows only"""
    return np.where(np.asarray(clf_prob).reshape(-1) > threshold)[0]


def zero_inflated_combine(clf_prob, reg_mean, threshold=0.5):
    """(probability-weighted, hard-cut) predictions of the two-stage zero-inflated baseline"""
    clf_prob = np.asarray(clf_prob, dtype=float)
    reg_mean = np.asarray(reg_mean, dtype=float)
    return clf_prob * reg_mean, (clf_prob > threshold) * 1.0 * reg_mean
