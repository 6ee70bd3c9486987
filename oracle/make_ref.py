"""Stage the reference's own model classes under oracle/_ref/ so that they travel to the GPU box (oracle/_ref is
git-ignored, not gpurun-ignored): ``bench.py --impl reference`` then times the UNMODIFIED reference head
(``kind: "reference"``) instead of the oracle port.  TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py): nothing
under protoasnet_b200/ imports this, and the staged files never enter the repository history.

Only runs where /root/reference exists (the build container).  Staged: the model package the head classes need to import
(src/models/{Video_XProtoNet,XProtoNet,ProtoPNet,resnet_features,densenet_features,vgg_features}.py) and
src/utils/receptive_field.py, byte for byte.
"""
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["src/models/Video_XProtoNet.py", "src/models/XProtoNet.py", "src/models/ProtoPNet.py",
         "src/models/resnet_features.py", "src/models/densenet_features.py", "src/models/vgg_features.py",
         "src/utils/receptive_field.py"]


def stage() -> bool:
    if not os.path.isdir(REF):
        return os.path.isdir(os.path.join(DST, "src", "models"))
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), d)
    return True


def load_reference_classes():
    """-> (Video_XProtoNet, XProtoNet) imported from oracle/_ref, or None when it has not been staged."""
    import sys
    if not os.path.isdir(os.path.join(DST, "src", "models")):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from src.models.Video_XProtoNet import Video_XProtoNet
    from src.models.XProtoNet import XProtoNet
    return Video_XProtoNet, XProtoNet


if __name__ == "__main__":
    print("staged" if stage() else "reference checkout not found; nothing staged")
