"""TEST INFRASTRUCTURE: a CPU evaluator with the CoalitionEngine.evaluate interface, backed by
oracle/restate.py, so the host logic (Game memo, sharding, estimators) can be tested without a GPU."""
from oracle import restate


class OracleEvaluator:
    supports_image_range = True     # Game may hand it a slice of the validation set (dist.sharded_evaluate)

    def __init__(self, cfg, w0, deltas, images, labels):
        self.cfg, self.w0, self.deltas, self.images, self.labels = cfg, w0, deltas, images, labels
        self.n_val = images.shape[0]
        self.calls = []

    def evaluate(self, rows, image_range=None):
        self.calls.append(len(rows))
        self.ranges = getattr(self, "ranges", []) + [image_range]
        lo, hi = image_range if image_range is not None else (0, self.n_val)
        correct, loss = [], []
        for row in rows:
            members = [j for j, r in enumerate(row) if r != 0]
            agg = restate.get_aggregated_model([self.deltas[j] for j in members], [row[j] for j in members])
            sd = restate.model_agg_lazy(self.w0, [agg] if agg is not None else [])
            if hi <= lo:
                correct.append(0)
                loss.append(0.0)
                continue
            _, _, det = restate.evaluation(sd, self.cfg, self.images[lo:hi], self.labels[lo:hi], return_details=True)
            correct.append(int(det["correct"]))
            loss.append(float(det["loss_sum"]))
        return correct, loss
