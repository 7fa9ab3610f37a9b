# smoke of the build with the stricter validator (host-side change in vk_scene_upload's checks)
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
