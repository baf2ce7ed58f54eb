from photon_b200.eval.icl import EvalGauntlet, ICLEvaluator, expand_task
