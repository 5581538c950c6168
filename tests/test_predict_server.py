"""Wire protocol of the batch-prediction server (predict_server.predict_batch_process) against the reference's own
process body (env/workflow.py:274-593): same queues, same message shapes, same acknowledgements.  CPU: the protocol with a
recording predictor; GPU: the served values against the reference's model."""
import queue
import threading
import time

import numpy as np
import pytest


class FakePredictor:
    def __init__(self):
        self.loaded, self.batches = [], []

    def load_reference_state_dict(self, sd):
        self.loaded.append(sd)

    def predict_batch(self, obs):
        self.batches.append(len(obs))
        tag = float(len(self.loaded))
        return (np.tile(np.arange(12, dtype=np.float32), (len(obs), 1)) + obs[:, :1].astype(np.float32),
                obs[:, 1].astype(np.float32) + 100 * tag)


def _serve(predictor, **extra):
    from texasholdempocker_rl_b200.predict_server import predict_batch_process
    q = {k: queue.Queue() for k in ("in", "update", "wf", "wf_ack", "train_model", "train_model_ack", "hold", "hold_ack", "out")}
    stop = threading.Event()
    result = {}

    def run():
        result["served"] = predict_batch_process(q["in"], {}, {}, None, {"model_param_dict": {}}, q["update"], q["wf"], q["wf_ack"], q["train_model"],
                                                 q["train_model_ack"], q["hold"], q["hold_ack"], q["out"], 3, "WARNING",
                                                 interrupt_callback=stop.is_set, predictor=predictor, **extra)
    t = threading.Thread(target=run, daemon=True)
    t.start()
    return q, stop, t, result


def test_protocol_matches_reference_process(reference_tree):
    from env.constants import TrainHoldStatus, WorkflowStatus
    fake = FakePredictor()
    q, stop, t, result = _serve(fake)
    try:
        # requests: (pid, observation) in, (pid, reward_value, action_probs) out -- env/workflow.py:428, 327
        for pid in range(40):
            q["in"].put((pid, np.full(93, pid, dtype=np.int32)))
        got = sorted((q["out"].get(timeout=5) for _ in range(40)), key=lambda m: m[0])
        assert [m[0] for m in got] == list(range(40))
        assert all(float(m[1]) == m[0] and m[2].shape == (12,) and m[2][0] == m[0] for m in got)
        assert sum(fake.batches) == 40 and len(fake.batches) < 40                        # waiting rows were batched
        # workflow status changes are acknowledged with the status itself (:409-413)
        q["wf"].put(WorkflowStatus.TRAINING)
        assert q["wf_ack"].get(timeout=5) == WorkflowStatus.TRAINING
        # model update: the state_dict is applied to the live predictor and acknowledged with the current status (:415-424)
        q["update"].put({"w": 1})
        assert q["wf_ack"].get(timeout=5) == WorkflowStatus.TRAINING and fake.loaded == [{"w": 1}]
        # training hand-shake (:440-462): hold -> WAITING_MODEL -> (step, weights) -> two acks, then the batch is served on the new weights
        q["hold"].put(17)
        q["in"].put((7, np.full(93, 7, dtype=np.int32)))
        assert q["hold_ack"].get(timeout=5) == TrainHoldStatus.WAITING_MODEL
        assert q["out"].empty()                                                            # the request waits for the weights
        q["train_model"].put((17, {"w": 2}))
        assert q["train_model_ack"].get(timeout=5) == WorkflowStatus.TRAINING
        assert q["hold_ack"].get(timeout=5) == (TrainHoldStatus.FINISHED_UPDATING, 17)
        pid, value, probs = q["out"].get(timeout=5)
        assert pid == 7 and float(value) == 7 + 200 and fake.loaded[-1] == {"w": 2}
        # outside TRAINING the pushed weights are skipped but the hand-shake still completes (:449-453)
        q["wf"].put(WorkflowStatus.EVALUATING)
        assert q["wf_ack"].get(timeout=5) == WorkflowStatus.EVALUATING
        q["hold"].put(18)
        q["in"].put((8, np.full(93, 8, dtype=np.int32)))
        assert q["hold_ack"].get(timeout=5) == TrainHoldStatus.WAITING_MODEL
        q["train_model"].put((18, {"w": 3}))
        assert q["train_model_ack"].get(timeout=5) == WorkflowStatus.TRAINING
        assert q["hold_ack"].get(timeout=5) == (TrainHoldStatus.FINISHED_UPDATING, 18)
        assert q["out"].get(timeout=5)[0] == 8 and fake.loaded[-1] == {"w": 2}
    finally:
        stop.set()
        t.join(timeout=5)
    assert result["served"] == 42


@pytest.mark.gpu
def test_server_serves_reference_model_values(reference_tree, phe):
    """Through the real BatchedPredictor: what MCTS.predict (rl/MCTS.py:380-408) would receive equals the reference's own model,
    before and after a weight update pushed through update_model_param_queue."""
    import torch
    from test_wave_mcts import _setup
    from env.constants import WorkflowStatus
    params, mp, model, env = _setup(reference_tree, 41)
    _, _, model2, _ = _setup(reference_tree, 42)
    net = phe.PolicyValueNet(**mp)
    net.load_reference_state_dict(model.state_dict())
    predictor = phe.BatchedPredictor(net, dtype=torch.float32, max_batch=64)
    q, stop, t, result = _serve(predictor, max_batch=64)
    obs = phe.policy_value.random_observations(50, seed=9)
    try:
        for which in (model, model2):
            if which is model2:
                q["update"].put(model2.state_dict())
                assert q["wf_ack"].get(timeout=20) == WorkflowStatus.DEFAULT
            for pid in range(50):
                q["in"].put((pid, obs[pid]))
            got = sorted((q["out"].get(timeout=20) for _ in range(50)), key=lambda m: m[0])
            with torch.no_grad():
                logits, value, _, _, _ = which(torch.from_numpy(obs))
            want_p, want_v = torch.softmax(logits, 1).numpy(), value.numpy()
            assert np.abs(np.stack([m[2] for m in got]) - want_p).max() < 1e-3
            assert np.abs(np.array([m[1] for m in got]) - want_v).max() < 2e-3
    finally:
        stop.set()
        t.join(timeout=10)
