"""TEST INFRASTRUCTURE: a CPU evaluator with the CoalitionEngine.evaluate interface, backed by
oracle/restate.py, so the host logic (Game memo, sharding, estimators) can be tested without a GPU."""
from oracle import restate


class OracleEvaluator:
    def __init__(self, cfg, w0, deltas, images, labels):
        self.cfg, self.w0, self.deltas, self.images, self.labels = cfg, w0, deltas, images, labels
        self.n_val = images.shape[0]
        self.calls = []

    def evaluate(self, rows):
        self.calls.append(len(rows))
        correct, loss = [], []
        for row in rows:
            members = [j for j, r in enumerate(row) if r != 0]
            agg = restate.get_aggregated_model([self.deltas[j] for j in members], [row[j] for j in members])
            sd = restate.model_agg_lazy(self.w0, [agg] if agg is not None else [])
            _, _, det = restate.evaluation(sd, self.cfg, self.images, self.labels, return_details=True)
            correct.append(int(det["correct"]))
            loss.append(float(det["loss_sum"]))
        return correct, loss
