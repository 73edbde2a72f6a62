"""The documents cite files of this repository; every cited path has to exist (guards against doc rot)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md")]
PREFIXES = ("crazyara_b200/", "tests/", "tools/", "oracle/", "profiles/", "include/")
GENERATED = ("oracle/_ref", "oracle/_build", "crazyara_b200/libara_b200.so", "crazyara_b200/ara_uci")


def test_cited_repository_paths_exist():
    missing = []
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc)).read()
        for m in re.finditer(r"`([^`\s]+)`", text):
            token = m.group(1)
            if not token.startswith(PREFIXES) or any(token.startswith(g) for g in GENERATED):
                continue
            path = token.split("::")[0].split(":")[0].rstrip(".,;)")
            if any(c in path for c in "*<>{}…") or path.endswith("/"):
                path = path.split("*")[0].split("<")[0].split("{")[0].split("…")[0]
                path = os.path.dirname(path) if not path.endswith("/") else path
            if path and not os.path.exists(os.path.join(ROOT, path)):
                missing.append((doc, token))
            elif "::" in token and path.endswith(".py"):      # a cited function / class has to be there as well
                name = token.split("::")[1].split("(")[0].rstrip(".,;)")
                if name and not re.search(rf"^\s*(def|class)\s+{re.escape(name)}\b", open(os.path.join(ROOT, path)).read(), re.M):
                    missing.append((doc, token))
    assert missing == []
