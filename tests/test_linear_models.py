"""the NaiveBayes and LogisticRegression the reference scripts also fit (kdd99.py:57-58,67; cicids17.py:61-62,71).
CPU tests: b200flow.linear (torch fp64, device-agnostic) against the numpy restatement in oracle/linear.py and against
scikit-learn as the independent pin.  GPU test: the pyspark shim's estimators call the same functions on device tensors."""
import numpy as np
import pytest
import torch

from oracle import linear as orc


def _flows(seed, n=4000, D=12, C=5):
    rng = np.random.default_rng(seed)
    y = rng.choice(C, size=n, p=np.array([0.5, 0.25, 0.15, 0.07, 0.03][:C]) / sum([0.5, 0.25, 0.15, 0.07, 0.03][:C]))
    centres = rng.uniform(0.5, 6.0, size=(C, D))
    x = rng.poisson(centres[y]).astype(np.float64)                       # nonnegative counts, like flow counters
    x[:, 3] = 7.0                                                        # a constant column (std 0)
    x[:, 5] *= 1000.0                                                    # a column on another scale
    return x, y.astype(np.int64)


def test_naive_bayes_equals_oracle_and_sklearn_theta():
    from b200flow import linear
    sk = pytest.importorskip("sklearn.naive_bayes")
    x, y = _flows(0)
    fit = linear.nb_fit(torch.from_numpy(x), torch.from_numpy(y), 5, smoothing=1.0)
    pi, theta = orc.nb_fit(x, y, 5, 1.0)
    assert np.allclose(fit.pi.numpy(), pi, rtol=0, atol=1e-12) and np.allclose(fit.theta.numpy(), theta, rtol=0, atol=1e-12)
    m = sk.MultinomialNB(alpha=1.0).fit(x, y)
    assert np.allclose(fit.theta.numpy(), m.feature_log_prob_, rtol=0, atol=1e-12)          # independent pin (theta)
    # MLlib smooths the prior too: pi_c = log(n_c + 1) - log(N + C)
    n_c = np.bincount(y, minlength=5)
    assert np.allclose(fit.pi.numpy(), np.log(n_c + 1.0) - np.log(len(y) + 5.0), atol=1e-12)
    raw = linear.nb_raw(fit, torch.from_numpy(x)).numpy()
    raw_o, prob_o, pred_o = orc.nb_predict(pi, theta, x)
    assert np.allclose(raw, raw_o, rtol=1e-13, atol=1e-9) and np.array_equal(raw.argmax(1).astype(np.float64), pred_o)
    assert np.allclose(torch.softmax(torch.from_numpy(raw), 1).numpy(), prob_o, atol=1e-12)


def test_naive_bayes_absent_label_and_negative_values():
    from b200flow import linear
    x, y = _flows(1, n=500)
    y[y == 2] = 0                                                        # label 2 never occurs in the training rows
    fit = linear.nb_fit(torch.from_numpy(x), torch.from_numpy(y), 5)
    pi, theta = orc.nb_fit(x, y, 5)
    assert np.isneginf(fit.pi[2].item()) and np.isneginf(pi[2])
    keep = [0, 1, 3, 4]
    assert np.allclose(fit.pi.numpy()[keep], pi[keep], atol=1e-12) and np.allclose(fit.theta.numpy(), theta, atol=1e-12)
    # L = 4 labels present in the prior's denominator
    assert abs(fit.pi[0].item() - (np.log((y == 0).sum() + 1.0) - np.log(len(y) + 4.0))) < 1e-12
    assert not (linear.nb_raw(fit, torch.from_numpy(x)).argmax(1) == 2).any()
    x[3, 1] = -1.0
    with pytest.raises(ValueError):
        linear.nb_fit(torch.from_numpy(x), torch.from_numpy(y), 5)


@pytest.mark.parametrize("reg,alpha", [(0.3, 0.8), (0.05, 0.5), (0.01, 0.0)])
def test_logistic_regression_reaches_the_elastic_net_minimiser(reg, alpha):
    """OWL-QN (product) vs plain proximal gradient (oracle) vs scikit-learn saga: one strictly convex objective, one minimiser."""
    from b200flow import linear
    x, y = _flows(2, n=3000, D=8, C=4)
    xs, inv = orc.standardize(x)
    fit = linear.lr_fit(torch.from_numpy(x), torch.from_numpy(y), 4, max_iter=400, reg_param=reg, elastic_net=alpha, tol=1e-14,
                        family="multinomial")
    hist = fit.objective_history
    assert all(b <= a + 1e-12 for a, b in zip(hist, hist[1:]))                               # monotone decrease
    B_std = fit.coef.numpy() / np.where(inv > 0, inv, 1.0)                                    # back to the standardised scale
    assert np.all(fit.coef.numpy()[:, 3] == 0.0)                                             # constant column: coefficient 0
    obj = orc.lr_objective(xs, y, B_std, fit.intercept.numpy(), reg, alpha)
    assert abs(obj - hist[-1]) < 1e-12                                                       # same objective, restated in numpy
    B_o, b_o = orc.lr_minimise_ista(xs, y, 4, reg, alpha, iters=30000)
    obj_o = orc.lr_objective(xs, y, B_o, b_o, reg, alpha)
    assert obj <= obj_o + 1e-9 and abs(obj - obj_o) < 1e-6
    assert np.abs(B_std - B_o).max() < 5e-3 and np.abs(fit.intercept.numpy() - b_o).max() < 5e-3
    if alpha > 0:
        assert np.array_equal(np.abs(B_std) > 1e-6, np.abs(B_o) > 1e-6) or reg < 0.1         # same sparsity pattern (strong l1)
    skl = pytest.importorskip("sklearn.linear_model")
    m = skl.LogisticRegression(solver="saga", l1_ratio=alpha, C=1.0 / (len(y) * reg), max_iter=5000, tol=1e-10, fit_intercept=True).fit(xs, y)
    b_s = m.intercept_ - m.intercept_.mean()
    obj_s = orc.lr_objective(xs, y, m.coef_, b_s, reg, alpha)
    assert obj <= obj_s + 1e-7 and abs(obj - obj_s) < 1e-5                                   # independent pin


def test_logistic_regression_script_settings_twenty_iterations():
    """maxIter=20, regParam=0.3, elasticNetParam=0.8 (kdd99.py:57): 20 OWL-QN iterations end within 2e-4 of the minimum."""
    from b200flow import linear
    x, y = _flows(3, n=5000, D=12, C=5)
    xs, inv = orc.standardize(x)
    fit20 = linear.lr_fit(torch.from_numpy(x), torch.from_numpy(y), 5, max_iter=20, reg_param=0.3, elastic_net=0.8, family="multinomial")
    fit = linear.lr_fit(torch.from_numpy(x), torch.from_numpy(y), 5, max_iter=500, reg_param=0.3, elastic_net=0.8, tol=1e-14, family="multinomial")
    assert fit20.iterations <= 20 and fit20.objective_history[-1] - fit.objective_history[-1] < 2e-4
    raw20, raw = linear.lr_raw(fit20, torch.from_numpy(x)), linear.lr_raw(fit, torch.from_numpy(x))
    assert (raw20.argmax(1) == raw.argmax(1)).double().mean().item() > 0.995


def test_logistic_regression_binomial_pivot():
    from b200flow import linear
    skl = pytest.importorskip("sklearn.linear_model")
    x, y = _flows(4, n=3000, D=6, C=2)
    xs, inv = orc.standardize(x)
    fit = linear.lr_fit(torch.from_numpy(x), torch.from_numpy(y), 2, max_iter=300, reg_param=0.02, elastic_net=0.5, tol=1e-14)
    assert fit.binomial and fit.coef.shape == (1, 6)
    m = skl.LogisticRegression(solver="saga", l1_ratio=0.5, C=1.0 / (len(y) * 0.02), max_iter=5000, tol=1e-10).fit(xs, y)
    B_std = fit.coef.numpy() / np.where(inv > 0, inv, 1.0)
    assert np.abs(B_std - m.coef_).max() < 2e-3 and abs(fit.intercept.item() - m.intercept_[0]) < 2e-3
    raw = linear.lr_raw(fit, torch.from_numpy(x))
    prob = linear.lr_probability(fit, raw)
    assert torch.allclose(raw[:, 0], -raw[:, 1]) and torch.allclose(prob.sum(1), torch.ones(len(y), dtype=torch.float64))
    assert np.abs(prob.numpy() - m.predict_proba(xs)).max() < 2e-3


@pytest.mark.gpu
def test_shim_estimators_use_these_functions_on_the_device():
    from b200flow import linear, synth
    from pyspark.ml import Pipeline
    from pyspark.ml.classification import LogisticRegression, NaiveBayes
    from pyspark.ml.feature import StringIndexer, VectorAssembler
    from pyspark.sql import DataFrame
    rec, dicts = synth.make_kdd(20000, 5, seed=31, device="cuda")
    df = DataFrame.fromRecords(rec, synth.kdd_schema(), dicts)
    cats = synth.KDD_CATEGORICAL
    df = Pipeline(stages=[StringIndexer(inputCol=c, outputCol=c + "_num") for c in cats + ["label"]]).fit(df).transform(df)
    numerical = [c for c in df.columns if c not in cats + ["label", "label_num"]]
    df = VectorAssembler(inputCols=numerical, outputCol="features").transform(df)
    x = df._cols["features"].data.to(torch.float64).cpu().numpy()
    y = df._column_tensor("label_num").cpu().numpy().astype(np.int64)
    C = int(y.max()) + 1
    nb = NaiveBayes(labelCol="label_num", featuresCol="features", smoothing=1.0, modelType="multinomial").fit(df)
    pi, theta = orc.nb_fit(x, y, C, 1.0)
    assert np.allclose(nb.pi, pi, atol=1e-10) and np.allclose(nb.theta, theta, atol=1e-10)
    out = nb.transform(df)
    raw_o, prob_o, pred_o = orc.nb_predict(pi, theta, x)
    agree = (out._cols["prediction"].data.cpu().numpy() == pred_o).mean()
    assert agree > 0.9999                                                 # fp64 sums in another order: ties only
    lr = LogisticRegression(maxIter=20, regParam=0.3, elasticNetParam=0.8, featuresCol="features", labelCol="label_num",
                            family="multinomial").fit(df)
    ref = linear.lr_fit(torch.from_numpy(x), torch.from_numpy(y), C, max_iter=20, reg_param=0.3, elastic_net=0.8, family="multinomial")
    assert np.allclose(lr.coefficientMatrix, ref.coef.numpy(), atol=1e-6) and np.allclose(lr.interceptVector, ref.intercept.numpy(), atol=1e-6)
    xs, inv = orc.standardize(x)
    obj = orc.lr_objective(xs, y, lr.coefficientMatrix / np.where(inv > 0, inv, 1.0), lr.interceptVector, 0.3, 0.8)
    assert abs(obj - lr.summary.objectiveHistory[-1]) < 1e-9
