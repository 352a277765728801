// main.cpp — `RayTracing [scene]`, same CLI as the reference (src/main.cpp:5-16).
#include <string>

#include "scene.h"

int main(int argc, char** argv)
{
    const std::string scenePath = argc > 1 ? argv[1] : "input/simple_shapes.scene";
    Scene(scenePath).render();
    return 0;
}
