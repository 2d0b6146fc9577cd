import sys, os
sys.path.insert(0, ".")
os.environ["AGB_VERBOSE"] = "1"
if len(sys.argv) > 1:
    import torch
    torch.cuda.set_device(0)
    print("torch loaded", torch.version.cuda)
import alphagomoku_b200 as agb
eng = agb.Engine(agb.GameConfig(agb.GameRules(0), 15, 15), max_boards=1024 * 8, blocks=2, filters=64, games=1024, max_batch_size=8, max_simulations=100,
                 solver_max_positions=100, use_symmetries=True)
print(eng.stats()["solver_sms"], eng.stats()["pipeline_groups"])
